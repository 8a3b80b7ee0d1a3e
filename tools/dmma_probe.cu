// Probe: FP64 pipe peak on sm_100a — DFMA vs DMMA shapes. Scratch tool, results in profiles/.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

template<int ILP> __global__ void k_dfma(double* out, int iters, double a, double b){
  double acc[ILP];
  #pragma unroll
  for(int i=0;i<ILP;i++) acc[i]=threadIdx.x*1e-3+i;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<ILP;i++) acc[i]=fma(acc[i],a,b);
  }
  double s=0; 
  #pragma unroll
  for(int i=0;i<ILP;i++) s+=acc[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

template<int ILP> __global__ void k_m8n8k4(double* out, int iters, double a, double b){
  double c0[ILP], c1[ILP];
  #pragma unroll
  for(int i=0;i<ILP;i++){c0[i]=threadIdx.x*1e-3+i; c1[i]=i;}
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<ILP;i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<ILP;i++) s+=c0[i]+c1[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int ILP> __global__ void k_m16n8k4(double* out, int iters, double a, double b){
  double c[ILP][4];
  #pragma unroll
  for(int i=0;i<ILP;i++){c[i][0]=threadIdx.x*1e-3+i; c[i][1]=i;c[i][2]=1;c[i][3]=2;}
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<ILP;i++)
      asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};" : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b), "d"(b));
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<ILP;i++) s+=c[i][0]+c[i][1]+c[i][2]+c[i][3];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int ILP> __global__ void k_m16n8k8(double* out, int iters, double a, double b){
  double c[ILP][4];
  #pragma unroll
  for(int i=0;i<ILP;i++){c[i][0]=threadIdx.x*1e-3+i; c[i][1]=i;c[i][2]=1;c[i][3]=2;}
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<ILP;i++)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a));
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<ILP;i++) s+=c[i][0]+c[i][1]+c[i][2]+c[i][3];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int ILP> __global__ void k_m16n8k16(double* out, int iters, double a, double b){
  double c[ILP][4];
  #pragma unroll
  for(int i=0;i<ILP;i++){c[i][0]=threadIdx.x*1e-3+i; c[i][1]=i;c[i][2]=1;c[i][3]=2;}
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<ILP;i++)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};" : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a), "d"(b), "d"(a));
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<ILP;i++) s+=c[i][0]+c[i][1]+c[i][2]+c[i][3];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

template<typename F> float timeit(F f){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); f(); cudaDeviceSynchronize();
  float best=1e30f;
  for(int r=0;r<5;r++){ cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best) best=ms; }
  return best;
}

int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  printf("dev %s sms %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  int sms=p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double)*sms*8*1024));
  int iters=20000;
  for(int tpb : {128,256,512,1024}){
    for(int bps : {1,2}){
      if(tpb*bps>2048) continue;
      int grid=sms*bps; double nthr=(double)grid*tpb; double nwarp=nthr/32;
      float ms;
      ms=timeit([&]{k_dfma<8><<<grid,tpb>>>(out,iters,1.0000001,1e-9);});
      printf("tpb %4d bps %d  DFMA ilp8      %8.2f TFLOP/s\n",tpb,bps, nthr*iters*8*2/ms*1e-9);
      ms=timeit([&]{k_m8n8k4<8><<<grid,tpb>>>(out,iters,1.0000001,1e-9);});
      printf("tpb %4d bps %d  m8n8k4 ilp8    %8.2f TFLOP/s\n",tpb,bps, nwarp*iters*8*(2.0*8*8*4)/ms*1e-9);
      ms=timeit([&]{k_m16n8k4<4><<<grid,tpb>>>(out,iters,1.0000001,1e-9);});
      printf("tpb %4d bps %d  m16n8k4 ilp4   %8.2f TFLOP/s\n",tpb,bps, nwarp*iters*4*(2.0*16*8*4)/ms*1e-9);
      ms=timeit([&]{k_m16n8k8<4><<<grid,tpb>>>(out,iters,1.0000001,1e-9);});
      printf("tpb %4d bps %d  m16n8k8 ilp4   %8.2f TFLOP/s\n",tpb,bps, nwarp*iters*4*(2.0*16*8*8)/ms*1e-9);
      ms=timeit([&]{k_m16n8k16<4><<<grid,tpb>>>(out,iters/2,1.0000001,1e-9);});
      printf("tpb %4d bps %d  m16n8k16 ilp4  %8.2f TFLOP/s\n",tpb,bps, nwarp*(iters/2)*4*(2.0*16*8*16)/ms*1e-9);
    }
  }
  // ILP sensitivity at 256 threads x 1 block/SM (8 warps/SM)
  {
    int tpb=256, grid=sms; double nwarp=(double)grid*tpb/32; float ms;
    ms=timeit([&]{k_m8n8k4<1><<<grid,tpb>>>(out,iters,1.0000001,1e-9);});
    printf("m8n8k4 ilp1 8warps  %8.2f TFLOP/s\n", nwarp*iters*1*(512.0)/ms*1e-9);
    ms=timeit([&]{k_m8n8k4<2><<<grid,tpb>>>(out,iters,1.0000001,1e-9);});
    printf("m8n8k4 ilp2 8warps  %8.2f TFLOP/s\n", nwarp*iters*2*(512.0)/ms*1e-9);
    ms=timeit([&]{k_m8n8k4<4><<<grid,tpb>>>(out,iters,1.0000001,1e-9);});
    printf("m8n8k4 ilp4 8warps  %8.2f TFLOP/s\n", nwarp*iters*4*(512.0)/ms*1e-9);
    ms=timeit([&]{k_m16n8k16<1><<<grid,tpb>>>(out,iters/2,1.0000001,1e-9);});
    printf("m16n8k16 ilp1 8warps  %8.2f TFLOP/s\n", nwarp*(iters/2)*1*(2.0*16*8*16)/ms*1e-9);
    ms=timeit([&]{k_m16n8k16<2><<<grid,tpb>>>(out,iters/2,1.0000001,1e-9);});
    printf("m16n8k16 ilp2 8warps  %8.2f TFLOP/s\n", nwarp*(iters/2)*2*(2.0*16*8*16)/ms*1e-9);
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
