// The __global__ kernels of one (sketch dtype, lanes-per-row) instantiation and the table of
// host launchers the C ABI (mmw_api.cu) calls them through.  Every instantiation is its own
// translation unit (mmw_inst.cu compiled with -DSIGSDP_T / -DSIGSDP_G / -DSIGSDP_NAME), so the
// eight of them build in parallel.
#pragma once
#include "mmw_device.cuh"

namespace sigsdp {

struct KernelSet {
    // raises the dynamic shared-memory limit of the staged kernels and reports how many blocks
    // of the fused kernel (row-sharded variant if `shard`) fit on one SM
    cudaError_t (*prepare)(size_t smem, int shard, int* blocks_per_sm);
    // n_iters MMW iterations in ONE persistent launch (cooperative: all blocks co-resident)
    cudaError_t (*fused)(const void* prob, int grid, size_t smem, int n_iters, int do_finish, int shard,
                         cudaStream_t st);
    // one kernel per phase (stepwise mode: profiling / debugging)
    void (*dual)(const void* prob, int grid, cudaStream_t st);
    void (*exp)(const void* prob, int grid, cudaStream_t st);   // after the dual kernel: next soft-max shift
    void (*loss)(const void* prob, int grid, int it_local, cudaStream_t st);
    void (*term)(const void* prob, int grid, size_t smem, const void* bin, void* bout, double coeff, double mu, int slot,
                 cudaStream_t st);
    void (*copy)(const void* prob, int grid, void* dst, cudaStream_t st);
    void (*gram)(const void* prob, int grid, size_t smem, cudaStream_t st);
    void (*record)(const void* prob, int it_local, int m_star, long long s, double a1, double mu, int nterms,
                   cudaStream_t st);
    // one thread block per independent instance
    cudaError_t (*batch)(const void* probs_dev, int count, size_t smem, int n_iters, unsigned long long seed,
                         int blocks_per_instance, cudaStream_t st);
    cudaError_t (*batch_occupancy)(size_t smem, int* blocks_per_sm);
};

#ifdef SIGSDP_T
template <typename T, int G>
__global__ void __launch_bounds__(NT, 2) k_fused(Prob<T> P, int n_iters, int do_finish) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ double sh[NWARP + 2];
    GridTeam team(&P.ctrl->bar);
    run_iterations<T, G>(P, team, n_iters, dyn_smem, sh, do_finish);
}
// row-sharded variant: the team spans the GPUs of the box (ShardTeam)
template <typename T, int G>
__global__ void __launch_bounds__(NT, 2) k_fused_rows(Prob<T> P, int n_iters) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ double sh[NWARP + 2];
    __shared__ double xr[8];
    __shared__ unsigned xh[MAXR * LLW];
    ShardTeam team(P, xr, xh);
    run_iterations<T, G>(P, team, n_iters, dyn_smem, sh, 0);
}
template <typename T, int G>
__global__ void __launch_bounds__(NT, 2) k_dual(Prob<T> P) {
    __shared__ double sh[NWARP + 2];
    phase_dual<T, G>(P, StepTeam(), *reinterpret_cast<const volatile double*>(&P.ctrl->smax_shift), sh);
}
// stepwise: what the fused kernel does between the dual barrier and the Gram phase -- this iteration's
// maximum becomes the next one's soft-max shift and the reduction key is cleared
template <typename T>
__global__ void k_after_dual(Ctrl* ctrl) {
    ctrl->smax_shift = dkey_any_inv(ld_u64(&ctrl->emax_key));
}
template <typename T, int G>
__global__ void __launch_bounds__(NT, 2) k_loss(Prob<T> P, int it_local) {
    __shared__ double sh[NWARP + 2];
    phase_loss<T, G>(P, StepTeam(), it_local, sh);
}
template <typename T, int G>
__global__ void __launch_bounds__(NT, 2) k_term(Prob<T> P, const T* bin, T* bout, double coeff, double mu, int slot) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ double sh[NWARP + 2];
    if (P.tl.enabled) {
        Stage<T> st;
        stage_setup(P, dyn_smem, st);
        if (G >= 8 && P.Dp == G * Vec<T>::N)
            phase_term_staged2<T, G>(P, StepTeam(), bin, bout, coeff, slot, sh, st);
        else
            phase_term_staged<T, G>(P, StepTeam(), bin, bout, coeff, slot, sh, st);
    } else {
        phase_term<T, G>(P, StepTeam(), bin, bout, coeff, mu, slot, sh);
    }
}
template <typename T, int G>
__global__ void __launch_bounds__(NT, 2) k_copy(Prob<T> P, T* dst) {
    phase_copy<T, G>(P, StepTeam(), dst);
}
template <typename T, int G>
__global__ void __launch_bounds__(NT, 2) k_gram(Prob<T> P) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ double sh[NWARP + 2];
    if (P.tl.enabled) {
        Stage<T> st;
        stage_setup(P, dyn_smem, st);
        if (G >= 8 && P.Dp == G * Vec<T>::N)
            phase_gram_staged2<T, G>(P, StepTeam(), sh, st);
        else
            phase_gram_staged<T, G>(P, StepTeam(), sh, st);
    } else {
        phase_gram<T, G>(P, StepTeam(), sh);
    }
}
// batch: one thread block per independent instance, __syncthreads as the team barrier
template <typename T, int G>
__global__ void __launch_bounds__(NT, 2) k_batch(const Prob<T>* probs, int n_iters, unsigned long long seed) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ double sh[NWARP + 2];
    __shared__ Prob<T> Ps;
    {
        const int* src = reinterpret_cast<const int*>(probs + blockIdx.x);
        int* dst = reinterpret_cast<int*>(&Ps);
        for (int i = threadIdx.x; i < (int)(sizeof(Prob<T>) / sizeof(int)); i += NT) dst[i] = src[i];
        __syncthreads();
        if (threadIdx.x == 0) {
            Ps.omega = nullptr;   // (the uploaded descriptor carries the instance id in its seed field)
            Ps.seed = seed + 0x9E3779B97F4A7C15ull * (Ps.seed + 1ull);
        }
        __syncthreads();
    }
    CtaTeam team;
    run_iterations<T, G>(Ps, team, n_iters, dyn_smem, sh);
}
// batch, B > 1 blocks per instance (cooperative launch): block b works for instance b / B
template <typename T, int G>
__global__ void __launch_bounds__(NT, 2) k_batch_team(const Prob<T>* probs, int n_iters, unsigned long long seed, int B) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ double sh[NWARP + 2];
    __shared__ Prob<T> Ps;
    {
        const int* src = reinterpret_cast<const int*>(probs + blockIdx.x / B);
        int* dst = reinterpret_cast<int*>(&Ps);
        for (int i = threadIdx.x; i < (int)(sizeof(Prob<T>) / sizeof(int)); i += NT) dst[i] = src[i];
        __syncthreads();
        if (threadIdx.x == 0) {
            Ps.omega = nullptr;
            Ps.seed = seed + 0x9E3779B97F4A7C15ull * (Ps.seed + 1ull);
        }
        __syncthreads();
    }
    BatchTeam team(&Ps.ctrl->bar, B, (int)(blockIdx.x % B));
    run_iterations<T, G>(Ps, team, n_iters, dyn_smem, sh);
}
template <typename T>
__global__ void k_record(Prob<T> P, int it_local, int m_star, long long s, double a1, double mu, int nterms) {
    TaylorState ts;
    ts.m_star = m_star;
    ts.s = s;
    ts.a1 = a1;
    ts.mu = mu;
    ts.c1 = 0.0;
    record_history(P, P.ctrl->iter + it_local, ts, nterms);
    P.ctrl->total_terms += nterms;
}

template <typename T, int G>
struct Launchers {
    static const Prob<T>& prob(const void* p) { return *static_cast<const Prob<T>*>(p); }
    static cudaError_t prepare(size_t smem, int shard, int* occ) {
        cudaError_t e = cudaSuccess;
        // the staged kernels are opted in to the device's whole shared memory once per device (a
        // per-solver limit would break an older solver with larger tiles on its next launch)
        static bool opted[64] = {false};
        int dev = 0;
        if ((e = cudaGetDevice(&dev))) return e;
        if (dev >= 0 && dev < 64 && !opted[dev]) {
            int optin = 0;
            if ((e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev))) return e;
            optin -= 4096;   // room for the kernels' static shared memory (reduction scratch, barrier words)
            if ((e = cudaFuncSetAttribute(k_fused<T, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin))) return e;
            if ((e = cudaFuncSetAttribute(k_fused_rows<T, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin))) return e;
            if ((e = cudaFuncSetAttribute(k_term<T, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin))) return e;
            if ((e = cudaFuncSetAttribute(k_gram<T, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin))) return e;
            if ((e = cudaFuncSetAttribute(k_batch<T, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin))) return e;
            if ((e = cudaFuncSetAttribute(k_batch_team<T, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin))) return e;
            opted[dev] = true;
        }
        if (shard) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k_fused_rows<T, G>, NT, smem);
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k_fused<T, G>, NT, smem);
    }
    static cudaError_t fused(const void* p, int grid, size_t smem, int n_iters, int do_finish, int shard, cudaStream_t st) {
        Prob<T> P = prob(p);
        if (shard) {
            void* args[] = {(void*)&P, (void*)&n_iters};
            return cudaLaunchCooperativeKernel((void*)k_fused_rows<T, G>, dim3(grid), dim3(NT), args, smem, st);
        }
        void* args[] = {(void*)&P, (void*)&n_iters, (void*)&do_finish};
        return cudaLaunchCooperativeKernel((void*)k_fused<T, G>, dim3(grid), dim3(NT), args, smem, st);
    }
    static void dual(const void* p, int grid, cudaStream_t st) { k_dual<T, G><<<grid, NT, 0, st>>>(prob(p)); }
    static void exp_(const void* p, int, cudaStream_t st) { k_after_dual<T><<<1, 1, 0, st>>>(prob(p).ctrl); }
    static void loss(const void* p, int grid, int it_local, cudaStream_t st) { k_loss<T, G><<<grid, NT, 0, st>>>(prob(p), it_local); }
    static void term(const void* p, int grid, size_t smem, const void* bin, void* bout, double coeff, double mu, int slot,
                     cudaStream_t st) {
        k_term<T, G><<<grid, NT, smem, st>>>(prob(p), static_cast<const T*>(bin), static_cast<T*>(bout), coeff, mu, slot);
    }
    static void copy(const void* p, int grid, void* dst, cudaStream_t st) { k_copy<T, G><<<grid, NT, 0, st>>>(prob(p), static_cast<T*>(dst)); }
    static void gram(const void* p, int grid, size_t smem, cudaStream_t st) { k_gram<T, G><<<grid, NT, smem, st>>>(prob(p)); }
    static void record(const void* p, int it_local, int m_star, long long s, double a1, double mu, int nterms, cudaStream_t st) {
        k_record<T><<<1, 1, 0, st>>>(prob(p), it_local, m_star, s, a1, mu, nterms);
    }
    // blocks_per_instance > 1: that many co-resident blocks share an instance (the caller has checked
    // that count * blocks_per_instance blocks fit on the device)
    static cudaError_t batch(const void* probs, int count, size_t smem, int n_iters, unsigned long long seed,
                             int blocks_per_instance, cudaStream_t st) {
        const Prob<T>* pp = static_cast<const Prob<T>*>(probs);
        if (blocks_per_instance <= 1) {
            k_batch<T, G><<<dim3((unsigned)count), dim3(NT), smem, st>>>(pp, n_iters, seed);
            return cudaGetLastError();
        }
        void* args[] = {(void*)&pp, (void*)&n_iters, (void*)&seed, (void*)&blocks_per_instance};
        return cudaLaunchCooperativeKernel((void*)k_batch_team<T, G>, dim3((unsigned)(count * blocks_per_instance)), dim3(NT), args, smem, st);
    }
    static cudaError_t batch_occupancy(size_t smem, int* occ) {
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k_batch_team<T, G>, NT, smem);
    }
};
#endif  // SIGSDP_T

// defined by the eight instantiation units
extern const KernelSet ks_f64_g4, ks_f64_g8, ks_f64_g16, ks_f64_g32;
extern const KernelSet ks_f32_g4, ks_f32_g8, ks_f32_g16, ks_f32_g32;

}  // namespace sigsdp
