/* rkfd_kernel_variant.cu - one (RKFD_BLOCK, RKFD_GSCR, RKFD_RIGID, RKFD_SPEC) instantiation of rkfd_step_kernel. */
#include "rkfd_kernel.cuh"

#ifndef RKFD_BLOCK
#error "compile with -DRKFD_BLOCK=.. -DRKFD_GSCR=.. -DRKFD_RIGID=.. -DRKFD_SPEC=.."
#endif
#define RKFD_CAT_(a,b,c,d,e) a##b##_##c##_##d##_##e
#define RKFD_CAT(a,b,c,d,e) RKFD_CAT_(a,b,c,d,e)

namespace rkfd {

static void launch(const StateDev &st, int cur, int mode, int nsteps, int grid, size_t smem, cudaStream_t stream)
{
  rkfd_step_kernel<RKFD_BLOCK, RKFD_GSCR != 0, RKFD_RIGID != 0, RKFD_SPEC><<<grid, RKFD_BLOCK, smem, stream>>>(st, cur, mode, nsteps);
}
static int blocks_per_sm(size_t smem)
{
  auto k = rkfd_step_kernel<RKFD_BLOCK, RKFD_GSCR != 0, RKFD_RIGID != 0, RKFD_SPEC>;
  if( cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ){ cudaGetLastError(); return -1; }
  int nb = 0;
  if( cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, RKFD_BLOCK, smem) != cudaSuccess ){ cudaGetLastError(); return -1; }
  return nb;
}
static int upload(const ModelDev *m, cudaStream_t stream)
{
  return (int)cudaMemcpyToSymbolAsync(c_model, m, sizeof(ModelDev), 0, cudaMemcpyHostToDevice, stream);
}
extern const KernelVariant RKFD_CAT(rkfd_variant_, RKFD_BLOCK, RKFD_GSCR, RKFD_RIGID, RKFD_SPEC) = { RKFD_BLOCK, RKFD_GSCR != 0, RKFD_RIGID != 0, RKFD_SPEC, launch, blocks_per_sm, upload };

}  // namespace rkfd
