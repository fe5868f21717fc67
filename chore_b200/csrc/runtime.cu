// Handle lifetime, error text, workspace and weight dispatch of libchore_b200.so.
#include "common.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

std::atomic<uint64_t> g_launch_count{0};

bool chore_pdl_enabled() {
    static const bool on = [] {
        const char *e = getenv("CHORE_B200_PDL");      // opt-in: measured no gain inside a CUDA graph (2.78 vs 2.69 ms per image)
        return e != nullptr && e[0] == '1';
    }();
    return on;
}
static thread_local char g_err[512] = "";

void chore_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int chore_dev_alloc(chore_handle *h, void **p, size_t bytes) {
    CHORE_CUDA(cudaMalloc(p, bytes ? bytes : 16));
    h->owned.push_back(*p);
    return CHORE_OK;
}

static int reserve(void **p, size_t *have, size_t bytes) {
    if (*have >= bytes) return CHORE_OK;
    if (*p) CHORE_CUDA(cudaFree(*p));   // cudaFree synchronises the device: pending users are done
    *p = nullptr;
    *have = 0;
    CHORE_CUDA(cudaMalloc(p, bytes));
    *have = bytes;
    return CHORE_OK;
}
int chore_ws_reserve(chore_handle *h, size_t bytes) { return reserve(&h->ws, &h->ws_bytes, bytes); }
int chore_ws2_reserve(chore_handle *h, size_t bytes) { return reserve(&h->ws2, &h->ws2_bytes, bytes); }
int chore_lbs_ws_reserve(chore_handle *h, size_t bytes) { return reserve(&h->lbs_ws, &h->lbs_ws_bytes, bytes); }

extern "C" const char *chore_last_error(void) { return g_err; }
extern "C" int chore_abi_version(void) { return 3; }
extern "C" uint64_t chore_launch_count(void) { return g_launch_count.load(); }

extern "C" int chore_create(int device, chore_handle **out) {
    CHORE_CHECK(out != nullptr, "out is NULL");
    *out = nullptr;
    int count = 0;
    CHORE_CUDA(cudaGetDeviceCount(&count));
    CHORE_CHECK(device >= 0 && device < count, "device %d out of range (%d visible)", device, count);
    cudaDeviceProp prop;
    CHORE_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        chore_set_error("device %d is sm_%d%d; libchore_b200 is built for sm_100a only", device, prop.major,
                        prop.minor);
        return CHORE_ERR_ARCH;
    }
    CHORE_CUDA(cudaSetDevice(device));
    chore_handle *h = new chore_handle();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    *out = h;
    return CHORE_OK;
}

extern "C" void chore_destroy(chore_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    encoder_plan_destroy(h);
    for (void *p : h->owned) cudaFree(p);
    if (h->ws) cudaFree(h->ws);
    if (h->ws2) cudaFree(h->ws2);
    if (h->lbs_ws) cudaFree(h->lbs_ws);
    if (h->bwd_ws) cudaFree(h->bwd_ws);
    delete h;
}

extern "C" int chore_load_weights(chore_handle *h, const chore_tensor_desc *tensors, int n) {
    CHORE_CHECK(h && tensors && n > 0, "bad arguments");
    CHORE_CUDA(cudaSetDevice(h->device));
    std::map<std::string, const chore_tensor_desc *> byname;
    for (int i = 0; i < n; ++i) {
        CHORE_CHECK(tensors[i].name && tensors[i].data && tensors[i].ndim >= 1 && tensors[i].ndim <= 4,
                    "tensor %d: bad descriptor", i);
        std::string name = tensors[i].name;
        if (name.rfind("module.", 0) == 0) name = name.substr(7);   // DDP prefix (generator.py:255-262)
        byname[name] = &tensors[i];
    }
    if (int rc = query_load_weights(h, byname)) return rc;
    if (int rc = encoder_load_weights(h, byname)) return rc;
    CHORE_CUDA(cudaDeviceSynchronize());
    return CHORE_OK;
}
