// Output collection over NVLink peer memory (frames are sharded, the only exchange is the all-gather of the packed
// output rows).  One process per GPU; every rank owns ONE exportable allocation [slot 0 | slot 1 | flags] and maps the other ranks'
// allocations through CUDA IPC.  A rank PUSHES its rows into every peer's slot with copy-engine transfers (cudaMemcpyAsync between
// peer-mapped pointers: no SM, no NCCL channel CTAs beside the issue-bound forward kernels) and then raises a flag word in the
// peer's memory; the receiver waits on its local flag words.  Flags are monotonically increasing step counters, so nothing is reset.
//
// Measured reason (profiles/r2_bench_8gpu.json): NCCL's all-gather kernel needs its ~16-32 CTAs to move 2.07 GB per rank and step
// (2 / 8 CTAs: 42.6 / 11.7 ms per step) and stretched the 8-GPU step from 4.26 to 5.38 ms.
#include "common.cuh"

namespace ach {

__global__ void peer_signal_kernel(unsigned* const* flags, int n, unsigned value) {
    const int i = threadIdx.x;
    if (i < n && flags[i]) {
        __threadfence_system();   // (the data was written by earlier copies of this stream; the fence orders this thread's view)
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags[i]), "r"(value) : "memory");
    }
}

// A peer that died never raises its flag: after PEER_WAIT_LIMIT_NS the kernel traps (sticky CUDA error on this context - the
// process fails loudly at its next synchronisation and the launcher tears the job down) instead of spinning forever on the GPU.
constexpr unsigned long long PEER_WAIT_LIMIT_NS = 60ull * 1000ull * 1000ull * 1000ull;

__global__ void peer_wait_kernel(const unsigned* flags, int n, unsigned value) {
    const int i = threadIdx.x;
    if (i < n) {
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        unsigned v;
        unsigned spins = 0;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
            if ((int)(v - value) >= 0) break;
            __nanosleep(200);
            if ((++spins & 0xfffu) == 0) {
                unsigned long long t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > PEER_WAIT_LIMIT_NS) {
                    printf("ach_peer_wait: flag %d still %u after 60 s (waiting for %u) - a peer rank is gone\n", i, v, value);
                    __trap();
                }
            }
        } while (true);
    }
}

}  // namespace ach

using namespace ach;

extern "C" int ach_peer_alloc(long long bytes, void** ptr) {
    ACH_REQUIRE(bytes > 0 && ptr, "ach_peer_alloc: bad args");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);   // plain cudaMalloc: exportable with cudaIpcGetMemHandle (pool / VMM memory is not)
    if (e == cudaSuccess) e = cudaMemset(p, 0, (size_t)bytes);
    if (e != cudaSuccess) {
        set_error("ach_peer_alloc: %s", cudaGetErrorString(e));
        if (p) cudaFree(p);
        return ACH_ERR_CUDA;
    }
    *ptr = p;
    return ACH_OK;
}

extern "C" int ach_peer_free(void* ptr) {
    const cudaError_t e = cudaFree(ptr);
    if (e != cudaSuccess) {
        set_error("ach_peer_free: %s", cudaGetErrorString(e));
        return ACH_ERR_CUDA;
    }
    return ACH_OK;
}

extern "C" int ach_peer_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

extern "C" int ach_peer_export(void* ptr, unsigned char* handle) {
    ACH_REQUIRE(ptr && handle, "ach_peer_export: null arg");
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
    if (e != cudaSuccess) {
        set_error("ach_peer_export: %s", cudaGetErrorString(e));
        return ACH_ERR_CUDA;
    }
    memcpy(handle, &h, sizeof(h));
    return ACH_OK;
}

extern "C" int ach_peer_open(const unsigned char* handle, void** ptr) {
    ACH_REQUIRE(ptr && handle, "ach_peer_open: null arg");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        set_error("ach_peer_open: %s", cudaGetErrorString(e));
        return ACH_ERR_CUDA;
    }
    *ptr = p;
    return ACH_OK;
}

extern "C" int ach_peer_close(void* ptr) {
    const cudaError_t e = cudaIpcCloseMemHandle(ptr);
    if (e != cudaSuccess) {
        set_error("ach_peer_close: %s", cudaGetErrorString(e));
        return ACH_ERR_CUDA;
    }
    return ACH_OK;
}

extern "C" int ach_peer_copy(void* dst, const void* src, long long bytes, void* stream) {
    ACH_REQUIRE(dst && src && bytes > 0, "ach_peer_copy: bad args");
    const cudaError_t e = cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) {
        set_error("ach_peer_copy: %s", cudaGetErrorString(e));
        return ACH_ERR_CUDA;
    }
    return ACH_OK;
}

extern "C" int ach_peer_signal(unsigned* const* flags, int n, unsigned value, void* stream) {
    ACH_REQUIRE(flags && n > 0 && n <= 64, "ach_peer_signal: bad args (n <= 64)");
    peer_signal_kernel<<<1, 64, 0, static_cast<cudaStream_t>(stream)>>>(flags, n, value);
    return check_launch("ach_peer_signal");
}

extern "C" int ach_peer_wait(const unsigned* flags, int n, unsigned value, void* stream) {
    ACH_REQUIRE(flags && n > 0 && n <= 64, "ach_peer_wait: bad args (n <= 64)");
    peer_wait_kernel<<<1, 64, 0, static_cast<cudaStream_t>(stream)>>>(flags, n, value);
    return check_launch("ach_peer_wait");
}
