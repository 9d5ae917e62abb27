// libxsb200 runtime: error string, device init, small utilities.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";
static int g_sm_count = 0;
static int g_device = -1;

void xsb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// SMs the grid-sized (persistent / one-wave) kernels may plan for. Under data parallelism NCCL's all-reduce CTAs run next to
// the backward kernels: a kernel that needs EVERY SM for its one CTA per SM then waits for the collective's CTAs to leave
// (the all-reduce is exposed although it "overlaps"); planning for 148 - k SMs leaves the collective its own.
static int g_sm_budget = -1;      // -1: read XSB_SM_BUDGET on first use; 0: all SMs

int xsb_sm_count_cached() {
    if (g_sm_budget < 0) {
        const char* e = getenv("XSB_SM_BUDGET");
        g_sm_budget = e ? atoi(e) : 0;
    }
    if (g_sm_count == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) {
            int n = 0;
            if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) g_sm_count = n;
        }
        if (g_sm_count == 0) g_sm_count = 148;
    }
    return (g_sm_budget > 0 && g_sm_budget < g_sm_count) ? g_sm_budget : g_sm_count;
}

extern "C" {

const char* xsb_last_error(void) { return g_err; }

int xsb_version(void) { return 100; }

// Bind the calling thread to `device` and verify it is a Blackwell (sm_100) part.
int xsb_init(int device) {
    XSB_CUDA(cudaSetDevice(device));
    int major = 0, minor = 0;
    XSB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    XSB_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
    if (major != 10) {
        xsb_set_error("libxsb200 is built for sm_100a only; device %d is sm_%d%d", device, major, minor);
        return XSB_ERR_UNSUPPORTED;
    }
    g_device = device;
    g_sm_count = 0;
    (void)xsb_sm_count_cached();
    return XSB_OK;
}

int xsb_device_sm_count(void) { return xsb_sm_count_cached(); }

// n > 0: grid-sized kernels plan for n SMs (see g_sm_budget); 0: all of them
int xsb_set_sm_budget(int n) {
    XSB_CHECK_ARG(n >= 0, "sm budget must be >= 0");
    g_sm_budget = n;
    return XSB_OK;
}

int xsb_stream_sync(void* stream) {
    XSB_CUDA(cudaStreamSynchronize(xsb_stream(stream)));
    return XSB_OK;
}

// Host <-> device copies for callers that hold plain host buffers (the e2e path of bench.py).
int xsb_memcpy_h2d(void* dst, const void* src, int64_t bytes, void* stream) {
    XSB_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, xsb_stream(stream)));
    return XSB_OK;
}

int xsb_memcpy_d2h(void* dst, const void* src, int64_t bytes, void* stream) {
    XSB_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, xsb_stream(stream)));
    return XSB_OK;
}

}  // extern "C"
