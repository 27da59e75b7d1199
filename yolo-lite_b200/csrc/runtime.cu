// yl11 runtime: error text, device binding, driver entry point for the TMA descriptor encoder.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace yl {

static thread_local char g_err[512] = "";
static EncodeTiledFn g_encode = nullptr;
static int g_device = -1;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return YL_ERR_CUDA;
}

EncodeTiledFn get_encode_tiled() { return g_encode; }

static int g_pdl_env = 1;                 // YL_PDL environment switch, read once in yl_init (default on)
static thread_local int t_pdl_user = 1;   // yl_set_pdl(): a launch attribute of the CALLING THREAD, like the current device

bool pdl_enabled() { return g_pdl_env != 0 && t_pdl_user != 0; }

int init_conv_tc();    // conv_tc.cu
int init_conv_chain(); // conv_chain.cu
int init_dwpw();       // dwpw_tc.cu
int init_nms();        // nms.cu
int init_attention();  // attention.cu
int init_pool();       // pool.cu
int init_c3k2();       // c3k2_fused.cu
int init_c3k2_tc();    // c3k2_tc.cu
int init_stem_fused(); // stem_fused.cu
int init_metrics();    // metrics.cu

}  // namespace yl

extern "C" {

int yl_version(void) { return YL11_VERSION; }

int yl_set_pdl(int enabled) {
    const int prev = yl::t_pdl_user;
    yl::t_pdl_user = enabled ? 1 : 0;
    return prev;
}

const char* yl_last_error_string(void) { return yl::g_err; }

int yl_init(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        yl::set_error("no CUDA device visible (%s); libyl11 has no CPU path", cudaGetErrorString(e));
        return YL_ERR_NO_DEVICE;
    }
    YL_CHECK(device >= 0 && device < count, YL_ERR_ARG, "device %d out of range (have %d)", device, count);
    YL_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    YL_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        yl::set_error("device %d is sm_%d%d; libyl11 is built for sm_100a only", device, prop.major, prop.minor);
        return YL_ERR_NO_DEVICE;
    }
    if (!yl::g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        YL_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        YL_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess, YL_ERR_CUDA,
                 "cuTensorMapEncodeTiled not available from the driver");
        yl::g_encode = reinterpret_cast<yl::EncodeTiledFn>(fn);
    }
    {
        const char* e = getenv("YL_PDL");
        yl::g_pdl_env = (e && *e) ? (atoi(e) != 0) : 1;
    }
    int rc;
    if ((rc = yl::init_conv_tc()) != 0) return rc;
    if ((rc = yl::init_conv_chain()) != 0) return rc;
    if ((rc = yl::init_dwpw()) != 0) return rc;
    if ((rc = yl::init_nms()) != 0) return rc;
    if ((rc = yl::init_attention()) != 0) return rc;
    if ((rc = yl::init_pool()) != 0) return rc;
    if ((rc = yl::init_c3k2()) != 0) return rc;
    if ((rc = yl::init_c3k2_tc()) != 0) return rc;
    if ((rc = yl::init_stem_fused()) != 0) return rc;
    if ((rc = yl::init_metrics()) != 0) return rc;
    yl::g_device = device;
    return YL_OK;
}

}  // extern "C"
