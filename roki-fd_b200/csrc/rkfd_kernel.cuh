/* rkfd_kernel.cuh - the fused step kernel template and its device context.  Each (BLOCK, GSCR, RIGID, SPEC)
 * variant is compiled in its own translation unit (rkfd_kernel_variant.cu with -D flags) so that the
 * variants build in parallel and the penalty-only kernels carry no rigid-solver code. */
#ifndef RKFD_KERNEL_CUH
#define RKFD_KERNEL_CUH

#include <cuda_runtime.h>

#include "rkfd_core.cuh"

namespace rkfd {

/* per-variant copy of the model table (each translation unit is its own module with its own constant bank) */
static __constant__ ModelDev c_model;

extern __shared__ double rkfd_smem[];

template <int BLOCK, bool GSCR, bool RIGID_>
struct DevCtx {
  static constexpr bool RIGID = RIGID_;
  StateDev st; int e, cur, tid;     /* e / tid: the SELECTED environment / scratch column (own, except in cooperative sections) */
  int e0, tid0, wsd;
  /* scratch element k of this thread: shared-memory column [k*BLOCK + tid] (LDS/STS, conflict-free) */
  __device__ __forceinline__ double &S(int k){
    if( GSCR ) return st.scratch[(size_t)k*st.ld + e];
    return rkfd_smem[k*BLOCK + tid];
  }
  /* warp-cooperative sections: all 32 lanes work on the environment of lane `src` */
  __device__ __forceinline__ int lanes() const { return 32; }
  __device__ __forceinline__ int lane() const { return tid0 & 31; }
  __device__ __forceinline__ unsigned ballot(bool p) const { return __ballot_sync(0xffffffffu, p); }
  __device__ __forceinline__ unsigned long long bcast(unsigned long long x, int src) const { return __shfl_sync(0xffffffffu, x, src); }
  __device__ __forceinline__ double allsum(double x) const {
#pragma unroll
    for(int o=16;o>0;o>>=1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x; }
  __device__ __forceinline__ void select(int src){ tid = (tid0 & ~31) | src; e = (e0 & ~31) | src; }
  __device__ __forceinline__ void unselect(){ tid = tid0; e = e0; }
  __device__ __forceinline__ void gsync() const { __syncwarp(); }
#ifndef RKFD_SYNC_LEVEL
#define RKFD_SYNC_LEVEL 2
#endif
  /* level 1: once per evaluation, 2: per pass, 3: per link iteration */
  __device__ __forceinline__ void phase_sync(int level) const { if( level <= RKFD_SYNC_LEVEL ) __syncthreads(); }
  __device__ __forceinline__ double &W(int i){ return st.ws[(size_t)(e0 >> 5)*wsd + i]; }
  /* per-env state in HBM: element k of the selected environment (global address space asserted: LDG/STG, not generic) */
  __device__ __forceinline__ double gld(const double *p, int k) const { const double *a = p + ((size_t)k*st.ld + e); __builtin_assume(__isGlobal(a)); return *a; }
  __device__ __forceinline__ void gst(double *p, int k, double v){ double *a = p + ((size_t)k*st.ld + e); __builtin_assume(__isGlobal(a)); *a = v; }
};

/* mode 0: nsteps x rkFDUpdate; 1: one non-committing evaluation; 2: one committing evaluation */
template <int BLOCK, bool GSCR, bool RIGID, int SPEC>
__global__ void __launch_bounds__(BLOCK) rkfd_step_kernel(StateDev st, int cur, int mode, int nsteps)
{
  const int e = blockIdx.x*BLOCK + threadIdx.x;
  if( e >= st.ld ) return;            /* whole warps only: the padding environments [B, ld) hold a valid zero state */
  DevCtx<BLOCK,GSCR,RIGID> ctx; ctx.st = st; ctx.e = ctx.e0 = e; ctx.cur = cur; ctx.tid = ctx.tid0 = threadIdx.x; ctx.wsd = c_model.ws_doubles;
  Core<DevCtx<BLOCK,GSCR,RIGID>, typename SpecOf<SPEC>::type> core(ctx);
  core.run(c_model, mode, nsteps);
}


/* one compiled variant: launch + occupancy query */
struct KernelVariant {
  int block; bool gscr, rigid; int spec;     /* spec: model specialisation id (rkfd_core.cuh), 0 = generic */
  void (*launch)(const StateDev &st, int cur, int mode, int nsteps, int grid, size_t smem, cudaStream_t stream);
  int (*blocks_per_sm)(size_t smem);      /* sets the dynamic shared memory attribute; <0 on error */
  int (*upload)(const ModelDev *m, cudaStream_t stream);   /* model table -> this variant's constant bank */
};

}  // namespace rkfd
#endif
