// Error reporting, device queries and small utility kernels.
#include <stdarg.h>
#include "internal.cuh"

namespace ipp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached = -1;
    if (cached < 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
        cached = n;
    }
    return cached;
}

// y[r] = sum_c A[r][c] x[c]; one warp per row, fixed summation order (lane-strided, then xor-shuffle tree)
__global__ void gemv_kernel(const double* __restrict__ A, int64_t lda, int64_t rows, int64_t cols,
                            const double* __restrict__ x, double* __restrict__ y) {
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const double* a = A + r * lda;
    double s = 0.0;
    for (int64_t c = lane; c < cols; c += 32) s = fma(a[c], x[c], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) y[r] = s;
}

int launch_gemv(const double* A, int64_t lda, int64_t rows, int64_t cols, const double* x, double* y, cudaStream_t s) {
    if (rows <= 0) return IPP_OK;
    const int wpb = 8;
    gemv_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, s>>>(A, lda, rows, cols, x, y);
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}

// Optional CUDA-event timing of the dominant kernel of ipp_gp_predict / ipp_gp_predict_i8 (the fused variance
// kernels), recorded on the launching stream.  Used by bench.py for the roofline figure; off by default.
constexpr int PROFILE_MAX = 8192;
static bool g_profile = false;
static cudaEvent_t g_ev[PROFILE_MAX][2];
static int g_ev_made = 0, g_ev_used = 0;

void profile_mark(int which, cudaStream_t s) {
    if (!g_profile || g_ev_used >= PROFILE_MAX) return;
    if (which == 0 && g_ev_used >= g_ev_made) {
        cudaEventCreate(&g_ev[g_ev_made][0]);
        cudaEventCreate(&g_ev[g_ev_made][1]);
        ++g_ev_made;
    }
    cudaEventRecord(g_ev[g_ev_used][which], s);
    if (which == 1) ++g_ev_used;
}

}  // namespace ipp

extern "C" int ipp_profile_enable(int on) {
    ipp::g_profile = on != 0;
    ipp::g_ev_used = 0;
    return IPP_OK;
}

extern "C" int ipp_profile_read(double* total_ms, int64_t* launches) {
    using namespace ipp;
    IPP_CHECK_ARG(total_ms && launches, "null pointer");
    double sum = 0.0;
    for (int i = 0; i < g_ev_used; ++i) {
        IPP_CUDA(cudaEventSynchronize(g_ev[i][1]));
        float ms = 0.f;
        IPP_CUDA(cudaEventElapsedTime(&ms, g_ev[i][0], g_ev[i][1]));
        sum += ms;
    }
    *total_ms = sum;
    *launches = g_ev_used;
    g_ev_used = 0;
    return IPP_OK;
}

extern "C" const char* ipp_last_error(void) { return ipp::g_err; }
extern "C" int ipp_version(void) { return 100; }
extern "C" int ipp_sm_count(void) { return ipp::sm_count(); }
