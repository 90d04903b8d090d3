// comm_vmm.cu -- exchange windows through the CUDA virtual-memory-management API.
//
// The fused all-reduce of the epilogue kernel (epilogue_common.cuh: peer_exchange) needs ONE small
// buffer per rank that all peers can read and write over NVLink.  Sharing it with
// cudaIpcOpenMemHandle(..., cudaIpcMemLazyEnablePeerAccess) switches peer access on for the whole
// device pair, after which the runtime treats EVERY later cudaMalloc / cudaFree of the process as
// peer-visible: measured ~20 ms per small allocation pair and 200-260 ms for a fresh 8 GB block
// with 3 peers (tools/exp_e2e_multi.py).  Here the window is a cuMemCreate allocation exported as
// a POSIX file descriptor; a peer imports it, maps it into its own address space and grants access
// to its own device only (cuMemSetAccess).  Nothing else of either process becomes peer-visible.
// The descriptors travel between the processes over a unix socket (multigpu.py, SCM_RIGHTS).
//
// The driver entry points are resolved with cudaGetDriverEntryPoint: no link-time dependency on
// libcuda.  Any failure is reported as FOS_ERR_UNSUPPORTED and the caller falls back to the
// cudaIpc windows (fos_comm_window_alloc / fos_comm_attach).
#include <cuda.h>
#include <string.h>
#include <unistd.h>

#include "fos_common.cuh"

namespace {

struct Drv {
    CUresult (*GetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*Create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*Release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*AddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*AddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*Map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*Unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*SetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*Export)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*Import)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
    bool ok = false;
};

template <typename F>
bool resolve(const char* name, F& fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess ||
        p == nullptr) {
        cudaGetLastError();
        return false;
    }
    fn = reinterpret_cast<F>(p);
    return true;
}

const Drv& drv() {
    static Drv d = [] {
        Drv x;
        x.ok = resolve("cuMemGetAllocationGranularity", x.GetAllocationGranularity) && resolve("cuMemCreate", x.Create) &&
               resolve("cuMemRelease", x.Release) && resolve("cuMemAddressReserve", x.AddressReserve) &&
               resolve("cuMemAddressFree", x.AddressFree) && resolve("cuMemMap", x.Map) && resolve("cuMemUnmap", x.Unmap) &&
               resolve("cuMemSetAccess", x.SetAccess) && resolve("cuMemExportToShareableHandle", x.Export) &&
               resolve("cuMemImportFromShareableHandle", x.Import);
        return x;
    }();
    return d;
}

#define FOS_DRV(expr)                                                                       \
    do {                                                                                    \
        CUresult _r = (expr);                                                               \
        if (_r != CUDA_SUCCESS) {                                                           \
            fos_set_error("%s:%d: %s -> CUresult %d", __FILE__, __LINE__, #expr, (int)_r);  \
            return FOS_ERR_UNSUPPORTED;                                                     \
        }                                                                                   \
    } while (0)

CUmemAllocationProp window_prop(int device) {
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof(prop));
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return prop;
}

int map_local(const Drv& d, CUmemGenericAllocationHandle handle, size_t size, size_t gran, int device, void** out) {
    CUdeviceptr p = 0;
    FOS_DRV(d.AddressReserve(&p, size, gran, 0, 0));
    CUresult r = d.Map(p, size, 0, handle, 0);
    if (r == CUDA_SUCCESS) {
        CUmemAccessDesc acc;
        memset(&acc, 0, sizeof(acc));
        acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        acc.location.id = device;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        r = d.SetAccess(p, size, &acc, 1);
        if (r != CUDA_SUCCESS) d.Unmap(p, size);
    }
    if (r != CUDA_SUCCESS) {
        d.AddressFree(p, size);
        fos_set_error("cannot map an exchange window into device %d (CUresult %d)", device, (int)r);
        return FOS_ERR_UNSUPPORTED;
    }
    *out = reinterpret_cast<void*>(p);
    return FOS_OK;
}

}  // namespace

// Release everything fos_comm_window_alloc_fd / fos_comm_attach_fd created (called by design_free).
void fos_comm_vmm_release(fos_design* h) {
    if (!h->vmm) return;
    const Drv& d = drv();
    for (int r = 0; r < FOS_MAX_WORLD; ++r) {
        if (h->vmm_ptr[r]) {
            d.Unmap(reinterpret_cast<CUdeviceptr>(h->vmm_ptr[r]), h->vmm_size);
            d.AddressFree(reinterpret_cast<CUdeviceptr>(h->vmm_ptr[r]), h->vmm_size);
            h->vmm_ptr[r] = nullptr;
        }
        if (h->vmm_handle[r]) {
            d.Release(static_cast<CUmemGenericAllocationHandle>(h->vmm_handle[r]));
            h->vmm_handle[r] = 0;
        }
    }
    if (h->vmm_fd >= 0) close(h->vmm_fd);
    h->vmm_fd = -1;
    h->vmm = false;
    h->window = nullptr;
}

extern "C" int fos_comm_window_alloc_fd(fos_design* h, int rank, int world, int* fd_out) {
    FOS_REQUIRE(h && fd_out, "null pointer argument");
    FOS_REQUIRE(world >= 1 && world <= FOS_MAX_WORLD, "world size must be in 1..%d", FOS_MAX_WORLD);
    FOS_REQUIRE(rank >= 0 && rank < world, "rank %d out of range", rank);
    FOS_REQUIRE(h->window == nullptr, "exchange window already allocated");
    FOS_CUDA(cudaSetDevice(h->device));
    FOS_CUDA(cudaFree(nullptr));  // make sure the primary context is current for the driver calls
    const Drv& d = drv();
    if (!d.ok) {
        fos_set_error("the CUDA driver does not expose the virtual-memory-management entry points");
        return FOS_ERR_UNSUPPORTED;
    }
    const CUmemAllocationProp prop = window_prop(h->device);
    size_t gran = 0;
    FOS_DRV(d.GetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
    h->window_bytes = fos_window_bytes(h);
    h->vmm_size = (h->window_bytes + gran - 1) / gran * gran;
    h->vmm_gran = gran;
    CUmemGenericAllocationHandle handle = 0;
    FOS_DRV(d.Create(&handle, h->vmm_size, &prop, 0));
    h->vmm = true;
    h->vmm_handle[rank] = handle;
    void* base = nullptr;
    int st = map_local(d, handle, h->vmm_size, gran, h->device, &base);
    int fd = -1;
    if (st == FOS_OK) {
        h->vmm_ptr[rank] = base;
        if (d.Export(&fd, handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) != CUDA_SUCCESS) {
            fos_set_error("cannot export the exchange window as a file descriptor");
            st = FOS_ERR_UNSUPPORTED;
        }
    }
    if (st != FOS_OK) {
        fos_comm_vmm_release(h);
        return st;
    }
    h->vmm_fd = fd;
    h->window = base;
    FOS_CUDA(cudaMemset(h->window, 0, h->vmm_size));
    h->rank = rank;
    h->world = 1;  // becomes `world` once the peers are attached
    fos_window_bind(h, rank, h->window);
    // the exchange counter starts at 1 so that a zeroed flag never satisfies a wait
    unsigned long long one = 1;
    FOS_CUDA(cudaMemcpy(h->peer.epoch, &one, sizeof(one), cudaMemcpyHostToDevice));
    *fd_out = fd;
    return FOS_OK;
}

// fds[r]: descriptor of rank r's window as received in THIS process (entry of the own rank ignored).
// The descriptors stay owned by the caller (close them after this call).
extern "C" int fos_comm_attach_fd(fos_design* h, const int* fds, int world) {
    FOS_REQUIRE(h && fds, "null pointer argument");
    FOS_REQUIRE(h->vmm && h->window != nullptr, "call fos_comm_window_alloc_fd first");
    FOS_REQUIRE(world >= 1 && world <= FOS_MAX_WORLD && h->rank < world, "bad world size %d", world);
    FOS_CUDA(cudaSetDevice(h->device));
    const Drv& d = drv();
    for (int r = 0; r < world; ++r) {
        if (r == h->rank) continue;
        CUmemGenericAllocationHandle handle = 0;
        CUresult cr = d.Import(&handle, reinterpret_cast<void*>(static_cast<uintptr_t>(fds[r])),
                               CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
        if (cr != CUDA_SUCCESS) {
            fos_set_error("cannot import the exchange window of rank %d (CUresult %d)", r, (int)cr);
            return FOS_ERR_COMM;
        }
        h->vmm_handle[r] = handle;
        void* base = nullptr;
        if (map_local(d, handle, h->vmm_size, h->vmm_gran, h->device, &base) != FOS_OK) return FOS_ERR_COMM;
        h->vmm_ptr[r] = base;
        fos_window_bind(h, r, base);
    }
    h->world = world;
    return fos_comm_after_attach(h);
}

// Drop a window that has not been attached yet (either kind), so that the other variant can be tried.
extern "C" int fos_comm_window_free(fos_design* h) {
    FOS_REQUIRE(h, "null design");
    FOS_REQUIRE(h->world == 1, "the window is attached to its peers; destroy the design instead");
    FOS_CUDA(cudaSetDevice(h->device));
    if (h->vmm) {
        fos_comm_vmm_release(h);
    } else if (h->window) {
        cudaFree(h->window);
        h->window = nullptr;
    }
    return FOS_OK;
}
