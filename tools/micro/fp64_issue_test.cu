// Tuning aid: does a DFMA occupy the warp scheduler's issue port for one cycle or two?  Per thread 8 independent DFMA
// chains and K independent IMAD chains per DFMA; if integer instructions issue in the shadow of the 2-cycle fp64 pipe,
// time stays flat up to K = 1; if the port is held, time grows as 2 + K.
#include <cstdio>
#include <cuda_runtime.h>
template <int K> __global__ void __launch_bounds__(256) k(double *out, int *iout, int iters, double a, double b, int ia, int ib)
{
  double x[8]; int y[16];
  for(int i=0;i<8;i++) x[i] = threadIdx.x*1e-3 + i;
  for(int i=0;i<16;i++) y[i] = threadIdx.x + i;
#pragma unroll 1
  for(int it=0; it<iters; it++){
#pragma unroll
    for(int r=0;r<4;r++){
#pragma unroll
      for(int i=0;i<8;i++){
        x[i] = fma(x[i], a, b);
#pragma unroll
        for(int j=0;j<K;j++) y[(2*i+j)&15] = y[(2*i+j)&15]*ia + ib;
      }
    }
  }
  double s = 0; int t = 0; for(int i=0;i<8;i++) s += x[i]; for(int i=0;i<16;i++) t += y[i];
  out[blockIdx.x*blockDim.x+threadIdx.x] = s; iout[blockIdx.x*blockDim.x+threadIdx.x] = t;
}
template <int K> void run(int sms){
  const int blocks = sms*8, threads = 256, iters = 2048;
  double *o; int *io; cudaMalloc(&o, blocks*threads*8); cudaMalloc(&io, blocks*threads*4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for(int rep=0;rep<4;rep++){ cudaEventRecord(e0); k<K><<<blocks,threads>>>(o, io, iters, 0.999999, 1e-7, 3, 7); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if( rep && ms < best ) best = ms; }
  const double dfma = 32.0*iters*blocks*threads;
  printf("K=%d IMAD per DFMA: %.3f ms, %.2f TFLOP/s fp64, %.3f cycles per DFMA warp-instruction per scheduler (1.965 GHz)\n", K, best, 2*dfma/best/1e9,
         best*1e-3*1.965e9/(dfma/32/(sms*4)));
  cudaFree(o); cudaFree(io);
}
int main(){ int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); run<0>(sms); run<1>(sms); run<2>(sms); run<3>(sms); return 0; }
