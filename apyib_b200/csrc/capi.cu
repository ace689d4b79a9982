// Library plumbing: error reporting, device info, CUDA-graph capture helpers and the
// calibration microbenchmarks that supply the FP64 roofline denominators.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace apyib {

static thread_local char g_err[512] = "no error";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- FP64 peak microbenchmarks -------------------------------------------------------
// Register-resident loops: 8 independent accumulator chains per warp so that the
// issue rate, not the dependent-issue latency, is what gets measured.
__global__ void __launch_bounds__(256) peak_dmma_kernel(double *sink, int iters) {
    double c[8][2];
#pragma unroll
    for (int q = 0; q < 8; ++q) c[q][0] = c[q][1] = 0.0;
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[q][0]), "+d"(c[q][1])
                         : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += c[q][0] + c[q][1];
    if (s == 123.456) sink[0] = s;
}

__global__ void __launch_bounds__(256) peak_dfma_kernel(double *sink, int iters) {
    double c[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) c[q] = 1e-3 * q;
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 16; ++q) c[q] = fma(c[q], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 16; ++q) s += c[q];
    if (s == 123.456) sink[0] = s;
}

__global__ void __launch_bounds__(256) peak_copy_kernel(double2 *dst, const double2 *src, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

}  // namespace apyib

using namespace apyib;

extern "C" int apyib_version(void) { return 100; }
extern "C" const char *apyib_last_error(void) { return g_err; }

extern "C" int apyib_device_info(int device, int *sm_count, int *cc_major, int *cc_minor, int64_t *mem_bytes) {
    cudaDeviceProp prop;
    APYIB_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (mem_bytes) *mem_bytes = (int64_t)prop.totalGlobalMem;
    return APYIB_OK;
}

extern "C" int apyib_graph_begin(void *stream) {
    APYIB_CUDA_CHECK(cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeThreadLocal));
    return APYIB_OK;
}

extern "C" int apyib_graph_end(void *stream, void **graph_exec_out) {
    APYIB_REQUIRE(graph_exec_out, "null pointer");
    cudaGraph_t graph = nullptr;
    APYIB_CUDA_CHECK(cudaStreamEndCapture((cudaStream_t)stream, &graph));
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    APYIB_CUDA_CHECK(e);
    *graph_exec_out = (void *)exec;
    return APYIB_OK;
}

extern "C" int apyib_graph_launch(void *graph_exec, void *stream) {
    APYIB_REQUIRE(graph_exec, "null graph");
    APYIB_CUDA_CHECK(cudaGraphLaunch((cudaGraphExec_t)graph_exec, (cudaStream_t)stream));
    return APYIB_OK;
}

extern "C" int apyib_graph_destroy(void *graph_exec) {
    if (graph_exec) APYIB_CUDA_CHECK(cudaGraphExecDestroy((cudaGraphExec_t)graph_exec));
    return APYIB_OK;
}

extern "C" int apyib_peak_fp64(int use_dmma, int iters, double *flops, float *ms_out) {
    APYIB_REQUIRE(flops && iters > 0, "arguments");
    int dev = 0, sms = 0;
    APYIB_CUDA_CHECK(cudaGetDevice(&dev));
    APYIB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double *sink = nullptr;
    APYIB_CUDA_CHECK(cudaMalloc(&sink, 64));
    cudaEvent_t e0, e1;
    APYIB_CUDA_CHECK(cudaEventCreate(&e0));
    APYIB_CUDA_CHECK(cudaEventCreate(&e1));
    const int blocks = sms * 8, threads = 256;
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        APYIB_CUDA_CHECK(cudaEventRecord(e0));
        if (use_dmma) peak_dmma_kernel<<<blocks, threads>>>(sink, iters);
        else peak_dfma_kernel<<<blocks, threads>>>(sink, iters);
        APYIB_CUDA_CHECK(cudaEventRecord(e1));
        APYIB_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0.f;
        APYIB_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    const double warps = (double)blocks * threads / 32.0;
    const double fl = use_dmma ? warps * iters * 8.0 * 512.0 : (double)blocks * threads * iters * 16.0 * 2.0;
    *flops = fl / (best * 1e-3);
    if (ms_out) *ms_out = best;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return APYIB_OK;
}

extern "C" int apyib_peak_copy(void *d_dst, const void *d_src, int64_t bytes, int iters, double *bytes_per_s) {
    APYIB_REQUIRE(d_dst && d_src && bytes_per_s && bytes >= 16 && iters > 0, "arguments");
    int dev = 0, sms = 0;
    APYIB_CUDA_CHECK(cudaGetDevice(&dev));
    APYIB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaEvent_t e0, e1;
    APYIB_CUDA_CHECK(cudaEventCreate(&e0));
    APYIB_CUDA_CHECK(cudaEventCreate(&e1));
    const int64_t n = bytes / 16;
    float best = 1e30f;
    for (int rep = 0; rep < iters + 1; ++rep) {
        APYIB_CUDA_CHECK(cudaEventRecord(e0));
        peak_copy_kernel<<<sms * 8, 256>>>((double2 *)d_dst, (const double2 *)d_src, n);
        APYIB_CUDA_CHECK(cudaEventRecord(e1));
        APYIB_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0.f;
        APYIB_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    *bytes_per_s = 2.0 * (double)n * 16.0 / (best * 1e-3);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return APYIB_OK;
}
