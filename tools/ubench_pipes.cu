// ubench_pipes.cu -- which issue pipe do the byte-scan building blocks of K1 use on sm_100a, and do two of them overlap?
// Each test runs 8 independent dependency chains per thread of ONE instruction kind (or two kinds, interleaved 1:1) and
// reports cycles per warp-instruction per SM sub-partition.  Two kinds that share a pipe add up; kinds on different
// pipes cost max(a, b) (until the 1 instr/clk issue limit).  Build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench tools/ubench_pipes.cu && /tmp/ubench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

#define ITERS 2048
#define CHAINS 8

#define OP_LOP3(x) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(k0), "r"(k1))
#define OP_IMAD(x) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(k0), "r"(k1))
#define OP_SHF(x) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(k0), "r"(k1))
#define OP_PRMT(x) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(k0), "r"(k1))
#define OP_IADD(x) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(k0))
#define OP_IADD3(x) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(x) : "r"(k0), "r"(k1))
#define OP_DP4A(x) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(x) : "r"(k0), "r"(k1))
#define OP_HSET2(x) asm volatile("set.ne.u32.f16x2 %0, %0, %1;" : "+r"(x) : "r"(k0))
#define OP_HADD2(x) asm volatile("add.f16x2 %0, %0, %1;" : "+r"(x) : "r"(k0))
#define OP_HFMA2(x) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x) : "r"(k0), "r"(k1))
#define OP_HMNMX2(x) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(x) : "r"(k0))
#define OP_FSET(x) asm volatile("set.ne.u32.f32 %0, %0, %1;" : "+r"(x) : "r"(k0))
#define OP_FFMA(x) asm volatile("{.reg .f32 a,b,c; mov.b32 a, %0; mov.b32 b, %1; mov.b32 c, %2; fma.rn.f32 a, a, b, c; mov.b32 %0, a;}" : "+r"(x) : "r"(k0), "r"(k1))
#define OP_IMNMX(x) asm volatile("min.u32 %0, %0, %1;" : "+r"(x) : "r"(k0))
#define OP_VOTE(x) asm volatile("{.reg .pred p; setp.ne.u32 p, %0, 0; vote.sync.ballot.b32 %0, p, 0xffffffff;}" : "+r"(x))
#define OP_SHFL(x) asm volatile("shfl.sync.down.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(x))
#define OP_POPC(x) asm volatile("popc.b32 %0, %0;" : "+r"(x))
#define OP_FFS(x) asm volatile("{.reg .u32 t; brev.b32 t, %0; bfind.shiftamt.u32 %0, t;}" : "+r"(x))
#define OP_IMADHI(x) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x) : "r"(k0))
#define OP_IMADW(x) asm volatile("{.reg .u64 t; mul.wide.u32 t, %0, %1; cvt.u32.u64 %0, t; }" : "+r"(x) : "r"(k0))
#define OP_SEL(x) asm volatile("{.reg .pred p; setp.ne.u32 p, %1, 0; selp.u32 %0, %0, %2, p;}" : "+r"(x) : "r"(k0), "r"(k1))
#define OP_ISETP(x) asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; @p add.u32 %0, %0, 1;}" : "+r"(x) : "r"(k0))
#define OP_VABSDIFF4(x) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(x) : "r"(k0), "r"(k1))
#define OP_LDS(x) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(sa) : "memory")
#define OP_LDS128(x) asm volatile("{.reg .u32 a,b,c; ld.shared.v4.u32 {%0,a,b,c}, [%1];}" : "=r"(x) : "r"(sa16) : "memory")
#define OP_LDS4WAY(x) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(sa4) : "memory")

#define DEFINE_SINGLE(NAME, OPA)                                                                  \
    __global__ void k_##NAME(uint32_t *out, uint32_t k0, uint32_t k1, long long *cyc) {           \
        __shared__ uint32_t sm[2048];                                                             \
        sm[threadIdx.x] = threadIdx.x;                                                            \
        __syncthreads();                                                                          \
        uint32_t sa = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 4;            \
        uint32_t sa16 = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 16;         \
        uint32_t sa4 = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 16 + 16;     \
        (void)sa; (void)sa16; (void)sa4;                                                          \
        uint32_t x[CHAINS];                                                                       \
        for (int i = 0; i < CHAINS; ++i) x[i] = threadIdx.x * 2654435761u + i;                    \
        long long t0 = clock64();                                                                 \
        for (int it = 0; it < ITERS; ++it) {                                                      \
            _Pragma("unroll") for (int i = 0; i < CHAINS; ++i) { OPA(x[i]); }                    \
        }                                                                                         \
        long long t1 = clock64();                                                                 \
        uint32_t acc = 0;                                                                         \
        for (int i = 0; i < CHAINS; ++i) acc ^= x[i];                                             \
        out[blockIdx.x * blockDim.x + threadIdx.x] = acc;                                         \
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                          \
    }

#define DEFINE_PAIR(NAME, OPA, OPB)                                                               \
    __global__ void k_##NAME(uint32_t *out, uint32_t k0, uint32_t k1, long long *cyc) {           \
        __shared__ uint32_t sm[2048];                                                             \
        sm[threadIdx.x] = threadIdx.x;                                                            \
        __syncthreads();                                                                          \
        uint32_t sa = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 4;            \
        uint32_t sa16 = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 16;         \
        uint32_t sa4 = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 16 + 16;     \
        (void)sa; (void)sa16; (void)sa4;                                                          \
        uint32_t x[CHAINS], y[CHAINS];                                                            \
        for (int i = 0; i < CHAINS; ++i) { x[i] = threadIdx.x * 2654435761u + i; y[i] = x[i] ^ 0x55u; } \
        long long t0 = clock64();                                                                 \
        for (int it = 0; it < ITERS; ++it) {                                                      \
            _Pragma("unroll") for (int i = 0; i < CHAINS; ++i) { OPA(x[i]); OPB(y[i]); }         \
        }                                                                                         \
        long long t1 = clock64();                                                                 \
        uint32_t acc = 0;                                                                         \
        for (int i = 0; i < CHAINS; ++i) acc ^= x[i] ^ y[i];                                      \
        out[blockIdx.x * blockDim.x + threadIdx.x] = acc;                                         \
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                          \
    }

DEFINE_SINGLE(lop3, OP_LOP3)
DEFINE_SINGLE(imad, OP_IMAD)
DEFINE_SINGLE(shf, OP_SHF)
DEFINE_SINGLE(prmt, OP_PRMT)
DEFINE_SINGLE(iadd, OP_IADD)
DEFINE_SINGLE(iadd3, OP_IADD3)
DEFINE_SINGLE(dp4a, OP_DP4A)
DEFINE_SINGLE(hset2, OP_HSET2)
DEFINE_SINGLE(hadd2, OP_HADD2)
DEFINE_SINGLE(hfma2, OP_HFMA2)
DEFINE_SINGLE(hmnmx2, OP_HMNMX2)
DEFINE_SINGLE(fset, OP_FSET)
DEFINE_SINGLE(ffma, OP_FFMA)
DEFINE_SINGLE(imnmx, OP_IMNMX)
DEFINE_SINGLE(vote, OP_VOTE)
DEFINE_SINGLE(shfl, OP_SHFL)
DEFINE_SINGLE(popc, OP_POPC)
DEFINE_SINGLE(ffs, OP_FFS)
DEFINE_SINGLE(imadhi, OP_IMADHI)
DEFINE_SINGLE(imadwide, OP_IMADW)
DEFINE_SINGLE(sel, OP_SEL)
DEFINE_SINGLE(isetp_padd, OP_ISETP)
DEFINE_SINGLE(vabsdiff4, OP_VABSDIFF4)
DEFINE_SINGLE(lds32, OP_LDS)
DEFINE_SINGLE(lds128, OP_LDS128)
DEFINE_SINGLE(lds32_4way, OP_LDS4WAY)
DEFINE_PAIR(lop3_imad, OP_LOP3, OP_IMAD)
DEFINE_PAIR(lop3_shf, OP_LOP3, OP_SHF)
DEFINE_PAIR(lop3_dp4a, OP_LOP3, OP_DP4A)
DEFINE_PAIR(imad_dp4a, OP_IMAD, OP_DP4A)
DEFINE_PAIR(lop3_hset2, OP_LOP3, OP_HSET2)
DEFINE_PAIR(imad_hset2, OP_IMAD, OP_HSET2)
DEFINE_PAIR(lop3_hadd2, OP_LOP3, OP_HADD2)
DEFINE_PAIR(imad_hadd2, OP_IMAD, OP_HADD2)
DEFINE_PAIR(lop3_hfma2, OP_LOP3, OP_HFMA2)
DEFINE_PAIR(lop3_hmnmx2, OP_LOP3, OP_HMNMX2)
DEFINE_PAIR(imad_hmnmx2, OP_IMAD, OP_HMNMX2)
DEFINE_PAIR(lop3_fset, OP_LOP3, OP_FSET)
DEFINE_PAIR(imad_fset, OP_IMAD, OP_FSET)
DEFINE_PAIR(lop3_ffma, OP_LOP3, OP_FFMA)
DEFINE_PAIR(lop3_vote, OP_LOP3, OP_VOTE)
DEFINE_PAIR(lop3_shfl, OP_LOP3, OP_SHFL)
DEFINE_PAIR(lop3_popc, OP_LOP3, OP_POPC)
DEFINE_PAIR(lop3_imadhi, OP_LOP3, OP_IMADHI)
DEFINE_PAIR(lop3_imadwide, OP_LOP3, OP_IMADW)
DEFINE_PAIR(lop3_lds128, OP_LOP3, OP_LDS128)
DEFINE_PAIR(lop3_vabsdiff4, OP_LOP3, OP_VABSDIFF4)
DEFINE_PAIR(imad_vabsdiff4, OP_IMAD, OP_VABSDIFF4)
DEFINE_PAIR(lop3_iadd, OP_LOP3, OP_IADD)

typedef void (*kern_t)(uint32_t *, uint32_t, uint32_t, long long *);
struct Test {
    const char *name;
    kern_t k;
    int ops_per_iter;  // warp-instructions per chain step as written (pairs: 2)
};

int main() {
    Test tests[] = {
        {"lop3", k_lop3, 1}, {"imad", k_imad, 1}, {"shf", k_shf, 1}, {"prmt", k_prmt, 1}, {"iadd", k_iadd, 1}, {"iadd_x2(iadd3?)", k_iadd3, 1},
        {"dp4a", k_dp4a, 1}, {"hset2", k_hset2, 1}, {"hadd2", k_hadd2, 1}, {"hfma2", k_hfma2, 1}, {"hmnmx2", k_hmnmx2, 1},
        {"fset", k_fset, 1}, {"ffma", k_ffma, 1}, {"imnmx", k_imnmx, 1}, {"setp+vote", k_vote, 1}, {"shfl", k_shfl, 1},
        {"popc", k_popc, 1}, {"brev+flo", k_ffs, 1}, {"imad.hi", k_imadhi, 1}, {"imad.wide", k_imadwide, 1}, {"setp+sel", k_sel, 1},
        {"setp+@p add", k_isetp_padd, 1}, {"vabsdiff4", k_vabsdiff4, 1}, {"lds32", k_lds32, 1}, {"lds128", k_lds128, 1},
        {"lds32 4-way conflict", k_lds32_4way, 1},
        {"lop3+imad", k_lop3_imad, 2}, {"lop3+shf", k_lop3_shf, 2}, {"lop3+dp4a", k_lop3_dp4a, 2}, {"imad+dp4a", k_imad_dp4a, 2},
        {"lop3+hset2", k_lop3_hset2, 2}, {"imad+hset2", k_imad_hset2, 2}, {"lop3+hadd2", k_lop3_hadd2, 2}, {"imad+hadd2", k_imad_hadd2, 2},
        {"lop3+hfma2", k_lop3_hfma2, 2}, {"lop3+hmnmx2", k_lop3_hmnmx2, 2}, {"imad+hmnmx2", k_imad_hmnmx2, 2},
        {"lop3+fset", k_lop3_fset, 2}, {"imad+fset", k_imad_fset, 2}, {"lop3+ffma", k_lop3_ffma, 2}, {"lop3+vote", k_lop3_vote, 2},
        {"lop3+shfl", k_lop3_shfl, 2}, {"lop3+popc", k_lop3_popc, 2}, {"lop3+imad.hi", k_lop3_imadhi, 2}, {"lop3+imad.wide", k_lop3_imadwide, 2},
        {"lop3+lds128", k_lop3_lds128, 2}, {"lop3+vabsdiff4", k_lop3_vabsdiff4, 2}, {"imad+vabsdiff4", k_imad_vabsdiff4, 2}, {"lop3+iadd", k_lop3_iadd, 2},
    };
    const int blocks = 148, threads = 512;  // 16 warps per SM = 4 per sub-partition
    uint32_t *d_out;
    long long *d_cyc, h_cyc[148];
    cudaMalloc(&d_out, blocks * threads * 4);
    cudaMalloc(&d_cyc, blocks * 8);
    printf("# cycles per written op per SM sub-partition (4 warps each, 8 chains per thread); pairs list cycles per PAIR\n");
    for (const Test &t : tests) {
        t.k<<<blocks, threads>>>(d_out, 3u, 5u, d_cyc);
        cudaDeviceSynchronize();
        t.k<<<blocks, threads>>>(d_out, 3u, 5u, d_cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("%-24s FAILED %s\n", t.name, cudaGetErrorString(e));
            return 1;
        }
        cudaMemcpy(h_cyc, d_cyc, blocks * 8, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < blocks; ++i) avg += (double)h_cyc[i];
        avg /= blocks;
        // per sub-partition: 4 warps x CHAINS x ITERS steps
        const double steps = 4.0 * CHAINS * ITERS;
        printf("%-24s %.3f cyc per %s\n", t.name, avg / steps, t.ops_per_iter == 2 ? "pair" : "op");
    }
    return 0;
}
