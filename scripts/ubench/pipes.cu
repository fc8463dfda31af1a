// Pipe-throughput microbenchmark for the integer instructions K1 is made of (sm_100a).
// One CTA of 1024 threads per SM; every thread runs ITER x 64 instructions over 8 independent
// chains; thread 0 reports clock64 cycles. rate = warp-instructions / cycle / SMSP.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define ITER 512
#define DECL uint32_t a0=s[0]+t,a1=s[1]+t,a2=s[2]+t,a3=s[3]+t,a4=s[4]+t,a5=s[5]+t,a6=s[6]+t,a7=s[7]+t; uint32_t b=s[8], c=s[9]|1;
#define FIN o[blockIdx.x*blockDim.x+threadIdx.x]=a0^a1^a2^a3^a4^a5^a6^a7;
#define R8(OP) OP(a0) OP(a1) OP(a2) OP(a3) OP(a4) OP(a5) OP(a6) OP(a7)
#define KERNEL(name, OP8)                                                          \
__global__ void __launch_bounds__(1024) name(const uint32_t* s, uint32_t* o, long long* cyc){ \
  uint32_t t=threadIdx.x; DECL                                                     \
  long long t0=clock64();                                                          \
  for(int i=0;i<ITER;++i){ OP8 OP8 OP8 OP8 OP8 OP8 OP8 OP8 }                       \
  long long t1=clock64(); FIN                                                      \
  if(threadIdx.x==0) cyc[blockIdx.x]=t1-t0; }

#define SHF(x) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
#define LOP(x) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c));
#define MIN2(x) asm volatile("min.u32 %0, %0, %1;" : "+r"(x) : "r"(b));
#define MIN3(x) asm volatile("{.reg .u32 q; min.u32 q, %0, %1; min.u32 %0, q, %2;}" : "+r"(x) : "r"(b), "r"(c));
#define MAD(x) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(c), "r"(b));
#define MADHI(x) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(c), "r"(b));
#define ADD(x) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(b));
#define ADD3(x) asm volatile("{.reg .u32 q; add.u32 q, %0, %1; add.u32 %0, q, %2;}" : "+r"(x) : "r"(b), "r"(c));
#define POPC(x) asm volatile("popc.b32 %0, %0;" : "+r"(x));
#define CLZ(x) asm volatile("clz.b32 %0, %0;" : "+r"(x));
#define SETSEL(x) asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %2, %0, p;}" : "+r"(x) : "r"(b), "r"(c));
#define PRMT(x) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
#define BFE(x) asm volatile("bfe.u32 %0, %0, 3, 9;" : "+r"(x));
#define SHLI(x) asm volatile("shl.b32 %0, %0, 3;" : "+r"(x));
#define SHRI(x) asm volatile("shr.u32 %0, %0, 3;" : "+r"(x));

KERNEL(k_shf, R8(SHF))
KERNEL(k_lop, R8(LOP))
KERNEL(k_min2, R8(MIN2))
KERNEL(k_min3, R8(MIN3))
KERNEL(k_mad, R8(MAD))
KERNEL(k_madhi, R8(MADHI))
KERNEL(k_add, R8(ADD))
KERNEL(k_add3, R8(ADD3))
KERNEL(k_popc, R8(POPC))
KERNEL(k_clz, R8(CLZ))
KERNEL(k_setsel, R8(SETSEL))
KERNEL(k_prmt, R8(PRMT))
KERNEL(k_bfe, R8(BFE))
KERNEL(k_shli, R8(SHLI))
KERNEL(k_shri, R8(SHRI))
// mixes: 4 of one + 4 of the other per group of 8
#define MIX(A,B) A(a0) B(a1) A(a2) B(a3) A(a4) B(a5) A(a6) B(a7)
KERNEL(k_shf_mad, MIX(SHF,MAD))
KERNEL(k_min3_mad, MIX(MIN3,MAD))
KERNEL(k_shf_lop, MIX(SHF,LOP))
KERNEL(k_shf_min3, MIX(SHF,MIN3))
KERNEL(k_lop_add, MIX(LOP,ADD))
KERNEL(k_shf_madhi, MIX(SHF,MADHI))
KERNEL(k_lop_mad, MIX(LOP,MAD))
KERNEL(k_min3_add, MIX(MIN3,ADD))
// 2 ALU : 1 FMA
#define MIX21(A,B) A(a0) A(a1) B(a2) A(a3) A(a4) B(a5) A(a6) A(a7)
KERNEL(k_shf2_mad1, MIX21(SHF,MAD))


#define SHFI(x) asm volatile("shf.l.wrap.b32 %0, %0, %1, 6;" : "+r"(x) : "r"(b));
#define SHFB(x) asm volatile("shf.r.wrap.b32 %0, %1, 0, %0;" : "+r"(x) : "r"(b));
#define LOPI(x) asm volatile("lop3.b32 %0, %0, %1, 0x2a, 0xea;" : "+r"(x) : "r"(b));
#define LOP2(x) asm volatile("and.b32 %0, %0, %1;" : "+r"(x) : "r"(b));
#define MINA(x) asm volatile("min.u32 %0, %0, %1;" : "+r"(x) : "r"(b));
#define MAXA(x) asm volatile("max.u32 %0, %0, %1;" : "+r"(x) : "r"(c));
#define MADI(x) asm volatile("mad.lo.u32 %0, %0, 5, %1;" : "+r"(x) : "r"(b));
#define ADDI(x) asm volatile("add.u32 %0, %0, 77;" : "+r"(x));
#define MIX4(A,B) A(a0) B(a0) A(a1) B(a1) A(a2) B(a2) A(a3) B(a3) A(a4) B(a4) A(a5) B(a5) A(a6) B(a6) A(a7) B(a7)
KERNEL(k_shfi, R8(SHFI))
KERNEL(k_shfb, R8(SHFB))
KERNEL(k_lopi, R8(LOPI))
KERNEL(k_lop2, R8(LOP2))
KERNEL(k_minmax, MIX(MINA,MAXA))
KERNEL(k_madi, R8(MADI))
KERNEL(k_shfi_minmax, MIX4(SHFI,MINA))
KERNEL(k_shfi_lop2, MIX4(SHFI,LOP2))
KERNEL(k_shfi_madi, MIX(SHFI,MADI))
KERNEL(k_lop2_madi, MIX4(LOP2,MADI))
KERNEL(k_lopi_min2, MIX4(LOPI,MINA))
KERNEL(k_shfb_min2, MIX4(SHFB,MINA))
KERNEL(k_shfi_min3, MIX(SHFI,MIN3))

// shared-memory LUT gather: random 4-byte reads from a 4 KB table
__global__ void __launch_bounds__(1024) k_lds(const uint32_t* s, uint32_t* o, long long* cyc){
  __shared__ uint32_t lut[1024];
  for(int i=threadIdx.x;i<1024;i+=blockDim.x) lut[i]=s[i&15]*2654435761u+i*40503u;
  __syncthreads();
  uint32_t t=threadIdx.x; DECL (void)b; (void)c;
  long long t0=clock64();
  for(int i=0;i<ITER;++i){
#pragma unroll
    for(int j=0;j<8;++j){
      a0=lut[a0&1023]; a1=lut[a1&1023]; a2=lut[a2&1023]; a3=lut[a3&1023];
      a4=lut[a4&1023]; a5=lut[a5&1023]; a6=lut[a6&1023]; a7=lut[a7&1023]; }
  }
  long long t1=clock64(); FIN
  if(threadIdx.x==0) cyc[blockIdx.x]=t1-t0; }

typedef void (*kfn)(const uint32_t*, uint32_t*, long long*);
struct Ent { const char* name; kfn f; int per; };
int main(){
  int dev=0; cudaSetDevice(dev); cudaDeviceProp p; cudaGetDeviceProperties(&p,dev);
  int nsm=p.multiProcessorCount;
  uint32_t hs[16]; for(int i=0;i<16;++i) hs[i]=0x9e3779b9u*(i+1);
  uint32_t *s,*o; long long *cyc; cudaMalloc(&s,64); cudaMalloc(&o,(size_t)nsm*1024*4); cudaMalloc(&cyc,nsm*8);
  cudaMemcpy(s,hs,64,cudaMemcpyHostToDevice);
  Ent es[]={{"SHF",k_shf,1},{"LOP3",k_lop,1},{"VIMNMX",k_min2,1},{"VIMNMX3",k_min3,1},{"IMAD",k_mad,1},{"IMAD.HI",k_madhi,1},
    {"ADD",k_add,1},{"ADD3",k_add3,1},{"POPC",k_popc,1},{"CLZ(FLO+IADD)",k_clz,1},{"SETP+SEL",k_setsel,1},{"PRMT",k_prmt,1},{"BFE",k_bfe,1},
    {"SHL imm",k_shli,1},{"SHR imm",k_shri,1},
    {"SHF+IMAD",k_shf_mad,1},{"VIMNMX3+IMAD",k_min3_mad,1},{"SHF+LOP3",k_shf_lop,1},{"SHF+VIMNMX3",k_shf_min3,1},{"LOP3+ADD",k_lop_add,1},
    {"SHF+IMAD.HI",k_shf_madhi,1},{"LOP3+IMAD",k_lop_mad,1},{"VIMNMX3+ADD",k_min3_add,1},{"2SHF+1IMAD",k_shf2_mad1,1},{"SHF imm (2reg)",k_shfi,1},{"SHF bit (2reg)",k_shfb,1},{"LOP3 2reg+imm",k_lopi,1},{"AND 2reg",k_lop2,1},{"IMAD r*imm+r",k_madi,1},{"SHFimm+MIN2",k_shfi_minmax,1},{"SHFimm+AND",k_shfi_lop2,1},{"SHFimm+IMADimm",k_shfi_madi,1},{"AND+IMADimm",k_lop2_madi,1},{"SHFimm+VIMNMX3",k_shfi_min3,1},{"LOPimm+MIN2",k_lopi_min2,1},{"SHFbit+MIN2",k_shfb_min2,1},{"LDS gather",k_lds,1}};
  for(auto&e:es){
    e.f<<<nsm,1024>>>(s,o,cyc); cudaDeviceSynchronize();
    e.f<<<nsm,1024>>>(s,o,cyc); cudaError_t er=cudaDeviceSynchronize();
    long long h[512]; cudaMemcpy(h,cyc,nsm*8,cudaMemcpyDeviceToHost);
    double avg=0; for(int i=0;i<nsm;++i) avg+=h[i]; avg/=nsm;
    double ops=(double)ITER*64*8; // PTX-level ops per warp x 8 warps per SMSP
    printf("%-16s cycles=%.0f  ptx-ops/clk/SMSP=%.3f  (%s)\n", e.name, avg, ops/avg, cudaGetErrorString(er));
  }
  return 0;
}
