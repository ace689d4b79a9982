"""Runtime switches of the drop-in layer."""
import os as _os

# Replay iterations 2.. of the CI solvers from one captured CUDA graph.
USE_CUDA_GRAPH = True
# The reference prints amplitude maxima / timings unconditionally (ci_wfn.py:132-133, 529-531;
# aats.py:552, 1053).  Kept for drop-in fidelity; benches and tests switch it off.
VERBOSE = True

# Device-resident mode (bench `value` leg): solver results stay on the GPU as torch tensors
# instead of being copied back to numpy; AAT accepts either.
RETURN_DEVICE = False
# Optional kernel timing hook: dict name -> list of (start_event, end_event); None = off.
TIMING = None
TIMING_ONLY = None          # optional name prefix filter for the hook


def timed(name):
    """Context manager recording CUDA events around a launch when TIMING is enabled."""
    import contextlib
    if TIMING is None or (TIMING_ONLY is not None and not name.startswith(TIMING_ONLY)):
        return contextlib.nullcontext()
    import torch
    if torch.cuda.is_current_stream_capturing():      # launches recorded into a CUDA graph carry no events
        return contextlib.nullcontext()

    @contextlib.contextmanager
    def _cm():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        yield
        e1.record()
        TIMING.setdefault(name, []).append((e0, e1))
    return _cm()

# How the substituted determinants of the AAT assembly are evaluated:
#   "lu"    : sub-warp LU with partial pivoting of every n x n matrix (csrc/dets.cu; the north star's
#             "batched small-LU/determinant kernel")
#   "lemma" : <= 4 x 4 determinants from S_oo^-1 (csrc/lemma.cu), same results
#   "factorized" : lemma for the small families, and the doubles x doubles table contracted in closed
#             form, O(o^2 v^2 (o+v)) DMMA contractions instead of (C(o,2) C(v,2))^2 determinants
AAT_ALGORITHM = "lu"

# Spatial CID / CISD (half-sum residual form, r2 = h + P h): evaluate the P-symmetric ladder term <ab|cd> t_ijcd over
# the o(o+1)/2 occupied pairs i <= j only (ci_wfn._PackedLadder) -- same result up to rounding, 46 % fewer ladder
# flops at o = 12.
PACKED_LADDER = _os.environ.get("APYIB_B200_PACKED_LADDER", "1") == "1"

# Batched solves (ci_wfn.solve_many): the finite-difference points are grouped by dtype (float64 / complex128) and,
# when a point's AO integrals exceed SOLVE_CHUNK_MIN_BYTES, split into chunks of at most SOLVE_CHUNK points in upload
# order; the groups are solved one after the other on the caller's stream while the copy stream keeps uploading the
# AO integrals of the later groups (complex points first, so their iterations hide the upload of the real points).
# SOLVE_CONCURRENT: run the groups on one CUDA stream each, driven from one host thread ("1"), never ("0"), or
# "auto" (default): only when every group is SMALL (points x o^2 v^2 <= SOLVE_CONCURRENT_MAX_ELEMS), i.e. when a
# group's ~40 dependent launches per iteration cannot fill the device on their own -- the H2O2/6-31G-sized
# molecules, and the 8-GPU runs of larger ones where a rank holds ~8 real points and one complex point.  Large
# groups saturate the device anyway (measured at N = 1, (S)-methyloxirane/cc-pVDZ shape: 1.06 s vs 1.02 s).
# The concurrent section runs WITHOUT the TMA-fed contraction kernel: TMA kernels running concurrently on different
# streams corrupt each other's loads (MO integrals off by 1e-2; bit-exact with USE_TMA = False or with
# CUDA_LAUNCH_BLOCKING=1; tools/diag_race3.py).  Passing the tensor maps through global memory (fence.proxy.tensormap)
# instead of as __grid_constant__ parameters was tried and does not change it; isolated, not fixed.
SOLVE_CONCURRENT = _os.environ.get("APYIB_B200_SOLVE_CONCURRENT", "auto")
SOLVE_CONCURRENT_MAX_ELEMS = 8_000_000
SOLVE_CHUNK = 32
SOLVE_CHUNK_MIN_BYTES = 64 << 20

# Route k-contiguous 2-D operand contractions (ladder term, first AO->MO quarter transform) through the
# TMA-fed kernel (csrc/contract_tma.cu); the gather kernel handles everything else.
USE_TMA = True

# LU path: feed the thread-per-matrix kernel column lists re-ordered for factorisation reuse (substituted
# columns last, lists sorted -> consecutive determinants share their leading panels).  Same results; False
# factorises every matrix from scratch (what bench.py's roofline line for the LU kernel is quoted on).
LU_REUSE = True

# LU path: evaluate the fused table x vector products of singly / doubly column-substituted tables with the
# prefix-shared LU kernel (csrc/dets_pairs.cu): one pivoted LU of the n-k unsubstituted columns per (row list,
# group), one Schur-complement k-vector per candidate column, k x k determinants per list.  Same quantity as the
# per-matrix LU (agreement ~1e-15 relative, not bitwise: the trailing k x k block is written out).
LU_PREFIX = _os.environ.get("APYIB_B200_LU_PREFIX", "1") == "1"

# Single-vector specialisation of the prefix-shared LU kernel for the doubles x doubles table -- the generic kernel
# carries predicated-off issue slots for up to 4 amplitude vectors in its pair loop (155 SASS instructions per pair
# against 31).  apyib_det_set_pairs_variant(1).  Validated on a B200 in round 2 (bit-identical to the generic kernel;
# 0.726 -> 0.552 ms per 12-overlap stack at n = 9): ON.
PAIRS_SINGLE_VECTOR = _os.environ.get("APYIB_B200_PAIRS_NY1", "1") == "1"

# OPTIONAL (validated on a B200 in round 2 against the host GEMV; host SCF is untimed, so off by default): (2J - K)[D] of the host SCF as one contraction launch per iteration
# on a device copy of 2(mn|ls) - (ml|ns) (hostchem._device_jk).
SCF_DEVICE_JK = _os.environ.get("APYIB_B200_SCF_DEVICE_JK", "0") == "1"

# AAT assembly: replay the device part of every overlap stack (aats.AAT._blocks_device) from a CUDA graph
# captured once per stack shape (static input buffers, private memory pool).  Same kernels, same order, same
# results; removes the host launch overhead of ~200 launches per stack on the LU path (H2O2/6-31G shape: 0.2045 -> 0.196 s
# per molecule; the step is then bound by the determinant kernels).
AAT_USE_GRAPH = _os.environ.get("APYIB_B200_AAT_GRAPH", "1") == "1"
