// Halo sums and scalar all-reduces of a sharded mesh through PEER MEMORY (NVLink / NVSwitch), without a host-launched
// collective (new functionality: the reference is single-GPU).
//
// Every rank owns one device region (cudaMalloc, exported with cudaIpcGetMemHandle, mapped by every other rank):
//   flags  [2][world]  uint64   epoch of the last completed push of rank r into this region (per buffer parity)
//   scal   [2][world][16] double rank r's partial scalars (energy, p.Hp, ...) of that epoch
//   recv   [2][rows]     rows pushed by the other ranks (their partial nodal sums on shared vertices: 3 nf scalars
//                        padded to whole 16-byte vectors, written with 16-byte stores)
// One exchange = two small kernels (128 threads, <= 40 registers, no shared memory: they co-reside with the persistent element
// kernel that is working on the interior tiles meanwhile):
//   PUSH  every shared row of this rank's partial results is stored straight into the sharers' recv buffers, the
//         partial scalars into every rank's scal slot; then (system-scope fences, last block) the epoch is published
//         in every peer's flag with a release store.
//   PULL  waits (acquire loads on the LOCAL flags) until every rank's push of this epoch has landed, sums the
//         partials of every shared vertex in ascending RANK order (own partial in its own position: all replicas
//         end up bit-identical, exactly like apl_halo_unpack) and the scalars in rank order.
// Buffers alternate with the epoch parity: rank A can only start pushing epoch e+2 after its pull of e+1, which needed
// B's push of e+1, which B issued after its pull of e -- so two buffers are race-free with no further handshake.
// The epoch counter lives in device memory and is advanced by the pull kernel, so a captured CUDA graph of
// (element kernels, push, pull) replays correctly.
#include <cuda_runtime.h>

#include <cstring>
#include <string>
#include <vector>

#include "common.h"

#define APL_XCHG_MAX_WORLD 16
#define APL_XCHG_NSCAL 16

struct apl_xchg {
    int world = 1, rank = 0, device = 0;
    int64_t max_rows = 0;            // capacity of one recv buffer, in rows (80 bytes each)
    size_t off_hdr = 0, off_scal = 0, off_recv = 0, recv_bytes = 0, total_bytes = 0;
    char* base = nullptr;            // own region
    char* peer[APL_XCHG_MAX_WORLD] = {};   // every rank's region in this process's address space (own included)
    bool opened[APL_XCHG_MAX_WORLD] = {};
    unsigned long long* d_epoch = nullptr;
    unsigned int* d_counters = nullptr;    // [0] push blocks done, [1] pull blocks done
    // plan (device arrays owned by the caller)
    int64_t n_send = 0, n_shared = 0;
    const int64_t* send_index = nullptr;
    const int32_t* send_peer = nullptr;
    const int64_t* send_row = nullptr;
    const int64_t* shared = nullptr;
    const int32_t* row_ptr = nullptr;
    const int64_t* src = nullptr;
};

namespace apl {
namespace {

struct XchgPeers {
    char* base[APL_XCHG_MAX_WORLD];
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Launches of a PNCG iteration are flag-guarded (a trial that is not needed is a no-op): the exchange that follows such
// a launch must be a no-op too.  Every rank holds bit-identical scalars, so all ranks skip together and the epoch
// simply does not advance.  skip == nullptr: unconditional.
struct XchgSkip {
    const double* scal = nullptr;   // the PNCG workspace scalars
    int a = -1, b = -1;             // skipped if scal[a] != 0 or scal[b + joff] != 0
    int dyn_j = 0;                  // joff = (int)scal[APL_S_J]; also added to the scalar slot exchanged
};
__device__ __forceinline__ bool xchg_skipped(const XchgSkip& sk, int& joff) {
    joff = 0;
    if (sk.scal == nullptr) return false;
    if (sk.dyn_j) joff = (int)__ldcg(sk.scal + APL_S_J);
    if (sk.a >= 0 && __ldcg(sk.scal + sk.a) != 0.0) return true;
    if (sk.b >= 0 && __ldcg(sk.scal + sk.b + joff) != 0.0) return true;
    return false;
}

// bytes of one row of the receive buffers: 3 nf scalars padded to whole 16-byte vectors (<= 80 for nf = 3 doubles)
template <typename T>
__host__ __device__ __forceinline__ size_t xchg_row_bytes(int nf) { return ((size_t)3 * nf * sizeof(T) + 15) / 16 * 16; }

template <typename T>
__global__ void __launch_bounds__(128) xchg_push_kernel(XchgPeers peers, int world, int rank, long long n_send,
                                                        const long long* __restrict__ idx, const int* __restrict__ peer,
                                                        const long long* __restrict__ row, int nf, const T* f0, const T* f1,
                                                        const T* f2, int ld, const void* scal_in, int n_scal, int scal_f64,
                                                        size_t off_scal, size_t off_recv, size_t recv_bytes,
                                                        const unsigned long long* epoch, unsigned int* counter,
                                                        XchgSkip sk) {
    int joff;
    if (xchg_skipped(sk, joff)) return;
    const unsigned long long e = *epoch + 1ull;
    const size_t par = (size_t)(e & 1ull);
    const T* f[3] = {f0, f1, f2};
    // One row = the 3 nf values of a shared vertex, padded to a multiple of 16 bytes and written with 16-byte stores:
    // scalar 4-byte stores over NVLink made this kernel 33 us for 55 k rows (tools/xchg_timing.py, run r2z).
    const size_t rs = xchg_row_bytes<T>(nf);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_send;
         i += (long long)gridDim.x * blockDim.x) {
        const long long v = idx[i];
        union { T t[12]; uint4 q[sizeof(T) == 4 ? 3 : 6]; } buf;
#pragma unroll
        for (int j = 0; j < 12; ++j) buf.t[j] = (T)0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (k < nf) {
                const T* r = f[k] + v * ld;
                buf.t[3 * k] = r[0];
                buf.t[3 * k + 1] = r[1];
                buf.t[3 * k + 2] = r[2];
            }
        }
        uint4* dst = reinterpret_cast<uint4*>(peers.base[peer[i]] + off_recv + par * recv_bytes + (size_t)row[i] * rs);
        const int nq = (int)(rs / 16);
#pragma unroll
        for (int c = 0; c < (sizeof(T) == 4 ? 3 : 6); ++c)
            if (c < nq) dst[c] = buf.q[c];
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < world * n_scal) {
        const int q = threadIdx.x / n_scal, k = threadIdx.x % n_scal;
        double* slot = reinterpret_cast<double*>(peers.base[q] + off_scal) + (par * world + rank) * APL_XCHG_NSCAL;
        slot[k] = scal_f64 ? reinterpret_cast<const double*>(scal_in)[joff + k]
                           : (double)reinterpret_cast<const T*>(scal_in)[joff + k];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(counter, 1u);
        if (done == gridDim.x - 1) {
            *counter = 0u;
            __threadfence_system();
            for (int q = 0; q < world; ++q)
                if (q != rank)
                    st_release_sys(reinterpret_cast<unsigned long long*>(peers.base[q]) + par * world + rank, e);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(128) xchg_pull_kernel(char* own, int world, int rank, long long n_shared,
                                                        const long long* __restrict__ shared,
                                                        const int* __restrict__ row_ptr, const long long* __restrict__ src,
                                                        int nf, T* f0, T* f1, T* f2, int ld, void* scal_out, int n_scal,
                                                        int scal_f64, size_t off_scal, size_t off_recv, size_t recv_bytes,
                                                        unsigned long long* epoch, unsigned int* counter, XchgSkip sk) {
    int joff;
    if (xchg_skipped(sk, joff)) return;
    const unsigned long long e = *epoch + 1ull;
    const size_t par = (size_t)(e & 1ull);
    if ((int)threadIdx.x < world && (int)threadIdx.x != rank) {
        const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(own) + par * world + threadIdx.x;
        while (ld_acquire_sys(flag) < e) __nanosleep(64);
    }
    __syncthreads();
    const char* recv = own + off_recv + par * recv_bytes;
    const size_t rs = xchg_row_bytes<T>(nf);
    T* f[3] = {f0, f1, f2};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_shared;
         i += (long long)gridDim.x * blockDim.x) {
        const long long v = shared[i];
        for (int k = 0; k < nf; ++k) {
            T* mine = f[k] + v * ld;
            T a0 = 0, a1 = 0, a2 = 0;
            for (int en = row_ptr[i]; en < row_ptr[i + 1]; ++en) {
                const long long j = src[en];   // -1: this rank's own partial, else a row of the receive buffer
                if (j < 0) {
                    a0 += mine[0]; a1 += mine[1]; a2 += mine[2];
                } else {
                    const T* r = reinterpret_cast<const T*>(recv + (size_t)j * rs) + 3 * k;
                    a0 += __ldcg(r); a1 += __ldcg(r + 1); a2 += __ldcg(r + 2);
                }
            }
            mine[0] = a0; mine[1] = a1; mine[2] = a2;
        }
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < n_scal) {
        const double* slots = reinterpret_cast<const double*>(own + off_scal) + par * world * APL_XCHG_NSCAL;
        double s = 0.0;
        for (int r = 0; r < world; ++r) s += __ldcg(slots + r * APL_XCHG_NSCAL + threadIdx.x);
        if (scal_f64) reinterpret_cast<double*>(scal_out)[joff + threadIdx.x] = s;
        else reinterpret_cast<T*>(scal_out)[joff + threadIdx.x] = (T)s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(counter, 1u);
        if (done == gridDim.x - 1) {
            *counter = 0u;
            *epoch = e;
        }
    }
}

int grid_rows(long long n) {
    long long g = (n + 127) / 128;
    if (g < 1) g = 1;
    if (g > 592) g = 592;   // 4 blocks of 128 threads per SM: one or two rows per thread at the benchmark's plane sizes (the
                            // kernels run AFTER the element pass by default; a grid of 128 blocks cost ~10 us per kernel)
    return (int)g;
}

}  // namespace
}  // namespace apl

using namespace apl;

extern "C" {

void apl_xchg_destroy(apl_xchg_t* x) {
    if (!x) return;
    cudaSetDevice(x->device);
    for (int r = 0; r < x->world; ++r)
        if (r != x->rank && x->opened[r]) cudaIpcCloseMemHandle(x->peer[r]);
    cudaFree(x->base);
    cudaFree(x->d_epoch);
    cudaFree(x->d_counters);
    delete x;
}

int apl_xchg_create(int world, int rank, int device, int64_t max_rows, apl_xchg_t** out) {
    if (!out) { set_error("apl_xchg_create: out is NULL"); return APL_ERR_INVALID; }
    *out = nullptr;
    if (world < 1 || world > APL_XCHG_MAX_WORLD || rank < 0 || rank >= world || max_rows < 0 || device < 0) {
        set_error("apl_xchg_create: bad arguments (1 <= world <= 16)");
        return APL_ERR_INVALID;
    }
    apl_xchg* x = new apl_xchg();
    x->world = world; x->rank = rank; x->device = device; x->max_rows = max_rows;
    // [flags][capacity header, 256 bytes][scalars][two receive buffers]
    x->off_hdr = ((size_t)2 * world * 8 + 255) / 256 * 256;
    x->off_scal = x->off_hdr + 256;
    x->off_recv = x->off_scal + ((size_t)2 * world * APL_XCHG_NSCAL * 8 + 255) / 256 * 256;
    x->recv_bytes = ((size_t)max_rows * 80 + 255) / 256 * 256;   // rows of up to 3 fields of doubles, padded to 80 bytes
    x->total_bytes = x->off_recv + 2 * x->recv_bytes + 256;
    auto fail = [&](const char* what, cudaError_t e) {
        set_error(std::string(what) + ": " + cudaGetErrorString(e));
        apl_xchg_destroy(x);
        return APL_ERR_CUDA;
    };
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail("cudaSetDevice", e);
    if ((e = cudaMalloc((void**)&x->base, x->total_bytes)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMemset(x->base, 0, x->total_bytes)) != cudaSuccess) return fail("cudaMemset", e);
    // the capacity is published in the region: peers address the second receive buffer with THEIR recv_bytes, so
    // apl_xchg_connect refuses regions of a different capacity instead of letting a push write out of bounds
    if ((e = cudaMemcpy(x->base + x->off_hdr, &x->max_rows, sizeof(int64_t), cudaMemcpyHostToDevice)) != cudaSuccess)
        return fail("cudaMemcpy", e);
    if ((e = cudaMalloc((void**)&x->d_epoch, 8)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMemset(x->d_epoch, 0, 8)) != cudaSuccess) return fail("cudaMemset", e);
    if ((e = cudaMalloc((void**)&x->d_counters, 8)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMemset(x->d_counters, 0, 8)) != cudaSuccess) return fail("cudaMemset", e);
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return fail("cudaDeviceSynchronize", e);
    x->peer[rank] = x->base;
    *out = x;
    return APL_OK;
}

int apl_xchg_ipc_handle(apl_xchg_t* x, void* handle64) {
    if (!x || !handle64) { set_error("apl_xchg_ipc_handle: NULL argument"); return APL_ERR_INVALID; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles are exchanged as 64 bytes");
    cudaIpcMemHandle_t h;
    APL_CUDA_CHECK(cudaSetDevice(x->device));
    APL_CUDA_CHECK(cudaIpcGetMemHandle(&h, x->base));
    memcpy(handle64, &h, 64);
    return APL_OK;
}

int apl_xchg_connect(apl_xchg_t* x, const void* handles) {
    if (!x || !handles) { set_error("apl_xchg_connect: NULL argument"); return APL_ERR_INVALID; }
    APL_CUDA_CHECK(cudaSetDevice(x->device));
    for (int r = 0; r < x->world; ++r) {
        if (r == x->rank || x->opened[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + 64 * (size_t)r, 64);
        void* p = nullptr;
        APL_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        x->peer[r] = (char*)p;
        x->opened[r] = true;
        int64_t cap = -1;
        APL_CUDA_CHECK(cudaMemcpy(&cap, x->peer[r] + x->off_hdr, sizeof(int64_t), cudaMemcpyDeviceToHost));
        if (cap != x->max_rows) {
            set_error("apl_xchg_connect: rank " + std::to_string(r) + " was created with max_recv_rows = " + std::to_string(cap) +
                      ", this rank with " + std::to_string(x->max_rows) + " (the capacity must be the same on every rank)");
            return APL_ERR_INVALID;
        }
    }
    return APL_OK;
}

int apl_xchg_set_plan(apl_xchg_t* x, int64_t n_send, const int64_t* send_index, const int32_t* send_peer,
                      const int64_t* send_row, int64_t n_shared, const int64_t* shared, const int32_t* row_ptr,
                      const int64_t* src) {
    if (!x || n_send < 0 || n_shared < 0 || (n_send > 0 && (!send_index || !send_peer || !send_row)) ||
        (n_shared > 0 && (!shared || !row_ptr || !src))) {
        set_error("apl_xchg_set_plan: bad arguments");
        return APL_ERR_INVALID;
    }
    x->n_send = n_send; x->send_index = send_index; x->send_peer = send_peer; x->send_row = send_row;
    x->n_shared = n_shared; x->shared = shared; x->row_ptr = row_ptr; x->src = src;
    return APL_OK;
}

static int xchg_check(apl_xchg_t* x, int dtype, int nf, const void* f0, int ld, int n_scal, const void* scal) {
    if (!x) { set_error("apl_xchg: NULL handle"); return APL_ERR_INVALID; }
    if ((dtype != APL_F32 && dtype != APL_F64) || nf < 0 || nf > 3 || (nf > 0 && !f0) || (ld != 3 && ld != 4) ||
        n_scal < 0 || n_scal > APL_XCHG_NSCAL || (n_scal > 0 && !scal)) {
        set_error("apl_xchg: bad arguments (up to 3 fields, up to 16 scalars)");
        return APL_ERR_INVALID;
    }
    for (int r = 0; r < x->world; ++r)
        if (!x->peer[r]) { set_error("apl_xchg: apl_xchg_connect has not mapped every peer yet"); return APL_ERR_STATE; }
    return APL_OK;
}

}  // extern "C"

namespace apl {
// Internal forms used by the PNCG driver (pncg.cu): scalars may be doubles whatever the field type, and the exchange
// can be guarded by the workspace's skip flags (skip_scal == nullptr: unconditional).
int xchg_push_ex(apl_xchg* x, int dtype, int nf, const void* f0, const void* f1, const void* f2, int ld, const void* scal,
                 int n_scal, int scal_f64, const double* skip_scal, int skip_a, int skip_b, int dyn_j, cudaStream_t s) {
    int rc = xchg_check(x, dtype, nf, f0, ld, n_scal, scal);
    if (rc != APL_OK) return rc;
    XchgPeers peers;
    for (int r = 0; r < APL_XCHG_MAX_WORLD; ++r) peers.base[r] = x->peer[r];
    XchgSkip sk;
    sk.scal = skip_scal; sk.a = skip_a; sk.b = skip_b; sk.dyn_j = dyn_j;
    const long long n = nf > 0 ? x->n_send : 0;
    const int grid = grid_rows(n);
    if (dtype == APL_F32)
        xchg_push_kernel<float><<<grid, 128, 0, s>>>(peers, x->world, x->rank, n, (const long long*)x->send_index,
                                                     x->send_peer, (const long long*)x->send_row, nf, (const float*)f0,
                                                     (const float*)f1, (const float*)f2, ld, scal, n_scal, scal_f64,
                                                     x->off_scal, x->off_recv, x->recv_bytes, x->d_epoch, x->d_counters, sk);
    else
        xchg_push_kernel<double><<<grid, 128, 0, s>>>(peers, x->world, x->rank, n, (const long long*)x->send_index,
                                                      x->send_peer, (const long long*)x->send_row, nf, (const double*)f0,
                                                      (const double*)f1, (const double*)f2, ld, scal, n_scal, scal_f64,
                                                      x->off_scal, x->off_recv, x->recv_bytes, x->d_epoch, x->d_counters,
                                                      sk);
    APL_CUDA_CHECK(cudaGetLastError());
    return APL_OK;
}

int xchg_pull_ex(apl_xchg* x, int dtype, int nf, void* f0, void* f1, void* f2, int ld, void* scal, int n_scal,
                 int scal_f64, const double* skip_scal, int skip_a, int skip_b, int dyn_j, cudaStream_t s) {
    int rc = xchg_check(x, dtype, nf, f0, ld, n_scal, scal);
    if (rc != APL_OK) return rc;
    XchgSkip sk;
    sk.scal = skip_scal; sk.a = skip_a; sk.b = skip_b; sk.dyn_j = dyn_j;
    const long long n = nf > 0 ? x->n_shared : 0;
    const int grid = grid_rows(n);
    if (dtype == APL_F32)
        xchg_pull_kernel<float><<<grid, 128, 0, s>>>(x->base, x->world, x->rank, n, (const long long*)x->shared, x->row_ptr,
                                                     (const long long*)x->src, nf, (float*)f0, (float*)f1, (float*)f2, ld,
                                                     scal, n_scal, scal_f64, x->off_scal, x->off_recv, x->recv_bytes,
                                                     x->d_epoch, x->d_counters + 1, sk);
    else
        xchg_pull_kernel<double><<<grid, 128, 0, s>>>(x->base, x->world, x->rank, n, (const long long*)x->shared,
                                                      x->row_ptr, (const long long*)x->src, nf, (double*)f0, (double*)f1,
                                                      (double*)f2, ld, scal, n_scal, scal_f64, x->off_scal, x->off_recv,
                                                      x->recv_bytes, x->d_epoch, x->d_counters + 1, sk);
    APL_CUDA_CHECK(cudaGetLastError());
    return APL_OK;
}
}  // namespace apl

extern "C" {

int apl_xchg_push(apl_xchg_t* x, int dtype, int nf, const void* f0, const void* f1, const void* f2, int ld,
                  const void* scal, int n_scal, void* stream) {
    return apl::xchg_push_ex(x, dtype, nf, f0, f1, f2, ld, scal, n_scal, 0, nullptr, -1, -1, 0, (cudaStream_t)stream);
}

int apl_xchg_pull(apl_xchg_t* x, int dtype, int nf, void* f0, void* f1, void* f2, int ld, void* scal, int n_scal,
                  void* stream) {
    return apl::xchg_pull_ex(x, dtype, nf, f0, f1, f2, ld, scal, n_scal, 0, nullptr, -1, -1, 0, (cudaStream_t)stream);
}

}  // extern "C"
