// Micro-test: tensor memory as per-thread scratch with several co-resident CTAs per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_scratch_test tmem_scratch_test.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k(int *bad, unsigned *info, int iters, size_t pad)
{
  extern __shared__ double sm[];
  __shared__ unsigned tmem_addr;
  constexpr unsigned TCOLS = 128*((BLOCK+127)/128);
  if( threadIdx.x < 32 ){
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(&tmem_addr)), "r"(TCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned warp = threadIdx.x >> 5;
  const unsigned tbase = tmem_addr + (((warp & 3u)*32u) << 16) + (warp >> 2)*128u;
  unsigned smid, wid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
  if( (threadIdx.x & 31) == 0 ){ info[(blockIdx.x*(BLOCK/32) + warp)*3] = smid; info[(blockIdx.x*(BLOCK/32) + warp)*3+1] = wid; info[(blockIdx.x*(BLOCK/32) + warp)*3+2] = tmem_addr; }
  if( threadIdx.x == 0 ){ const int c = atomicAdd((int*)&info[100000 + smid], 1) + 1; atomicMax((int*)&info[101000 + smid], c); atomicAdd((int*)&info[102000 + (tmem_addr & 0xffff)/32], 1); }
  int nbad = 0;
#pragma unroll 1
  for(int it=0; it<iters; it++){
#pragma unroll 1
    for(int kk=0;kk<56;kk++){
      const double v = (double)(blockIdx.x*BLOCK + threadIdx.x) + 0.001*kk + 1000000.0*it;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" :: "r"(tbase + 2u*kk), "r"(__double2loint(v)), "r"(__double2hiint(v)) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    sm[threadIdx.x] = it; __syncthreads();
#pragma unroll 1
    for(int kk=0;kk<56;kk++){
      const double v = (double)(blockIdx.x*BLOCK + threadIdx.x) + 0.001*kk + 1000000.0*it;
      unsigned lo, hi;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(tbase + 2u*kk));
      asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(lo), "+r"(hi) :: "memory");
      if( __hiloint2double((int)hi, (int)lo) != v ) nbad++;
    }
  }
  if( nbad ) atomicAdd(bad, nbad);
  __syncthreads();
  if( threadIdx.x == 0 ) atomicAdd((int*)&info[100000 + smid], -1);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if( threadIdx.x < 32 ) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_addr), "r"(TCOLS) : "memory");
}
template <int BLOCK> void run(int grid, size_t smem, int iters)
{
  int *bad; unsigned *info; cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4); cudaMalloc(&info, 110000*4 + grid*(BLOCK/32)*12); cudaMemset(info, 0, 110000*4 + grid*(BLOCK/32)*12);
  cudaFuncSetAttribute(k<BLOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k<BLOCK>, BLOCK, smem);
  k<BLOCK><<<grid, BLOCK, smem>>>(bad, info, iters, 0);
  cudaError_t e = cudaDeviceSynchronize();
  int h = -1; cudaMemcpy(&h, bad, 4, cudaMemcpyDeviceToHost);
  unsigned hi[24*3]; int n = grid*(BLOCK/32) < 24 ? grid*(BLOCK/32) : 24; cudaMemcpy(hi, info, n*12, cudaMemcpyDeviceToHost);
  printf("block %d grid %d smem %zu blocks/SM %d: %s, mismatches %d | (smid,warpid,taddr):", BLOCK, grid, smem, nb, cudaGetErrorString(e), h);
  for(int i=0;i<(n<8?n:8);i++) printf(" (%u,%u,%x)", hi[3*i], hi[3*i+1], hi[3*i+2]);
  printf("\n");
  { static unsigned hx[3000]; cudaMemcpy(hx, info + 101000, 2000*4, cudaMemcpyDeviceToHost); unsigned mx = 0; for(int i=0;i<1000;i++) if( hx[i] > mx ) mx = hx[i];
    printf("   max concurrent CTAs on one SM: %u; taddr column histogram (x32):", mx); for(int i=0;i<16;i++) printf(" %u", hx[1000+i]); printf("\n"); }
  cudaFree(bad); cudaFree(info);
}
int main()
{
  { cudaFuncAttributes a; cudaFuncGetAttributes(&a, k<128>);
    printf("k<128>: regs %d static smem %zu local %zu maxDyn %d\n", a.numRegs, a.sharedSizeBytes, a.localSizeBytes, a.maxDynamicSharedSizeBytes);
    for(size_t sm : {1024, 16384, 32768, 55296}){ int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k<128>, 128, sm); printf("  occupancy k<128> smem %zu: %d\n", sm, nb); }
    for(size_t sm : {1024, 55296}){ int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k<256>, 256, sm); printf("  occupancy k<256> smem %zu: %d\n", sm, nb); } }
  run<128>(4096, 1024, 200); run<128>(4096, 55296, 200); run<256>(2048, 16384, 200); run<256>(2048, 110592, 200); run<512>(2048, 16384, 200);
  run<128>(2, 55296, 10); run<128>(8, 55296, 10); run<128>(148*4, 55296, 200); run<128>(2048, 55296, 200); run<128>(2048, 75000, 200);
  run<256>(1024, 110592, 200); run<128>(2048, 110592, 200);
  return 0;
}
