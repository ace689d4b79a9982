#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
s=$(date +%s); timeout 900 python -m pytest tests -m gpu -x -q -n 4 > gpurun_out/last_tests.log 2>&1; echo "tests rc=$? $(( $(date +%s) - s ))s"; tail -4 gpurun_out/last_tests.log
grep -n "FAILED\|Error" gpurun_out/last_tests.log | head -5
timeout 300 python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value', d['value'], 'e2e', d['e2e']['value'], d['phases_max_over_ranks'])"
