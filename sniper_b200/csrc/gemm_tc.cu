// tcgen05 tensor-core contraction kernel for sm_100a: GEMM / implicit-GEMM convolution (NHWC) /
// weight-gradient, operands staged global->shared by TMA (cp.async.bulk.tensor, 128B swizzle),
// tcgen05.mma issued by one elected thread, fp32 accumulators in TMEM, epilogue via tcgen05.ld.
//
// Replaces the reference's library calls on the hot path: cudnnConvolutionForward/BackwardData/
// BackwardFilter (SNIPER-mxnet/src/operator/nn/cudnn/cudnn_convolution-inl.h:144,211-266), cuBLAS
// via FullyConnected (nn/fully_connected-inl.h) and linalg_gemm in DeformableConvolution
// (contrib/deformable_convolution-inl.h:148-160).
//
// One CTA = one 128 x BLOCK_N output tile (cta_group::1).  Warp roles: warp 0 = TMA producer,
// warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogue (TMEM lane quarter x column half each).
//
//   mode 0  GEMM    C[M,N]  = A[M,K] * B[N,K]^T             A,B K-major
//   mode 1  CONV    C[pix,N] = sum_taps A[n,h+dh,w+dw,c] * B[N,(tap,c)]   A 4-D NHWC, K-major; TMA zero
//                   fill implements padding, elementStrides implement stride
//   mode 2  WGRAD   C[Co,(tap,ci)] += sum_pix dY[pix,Co] * X[n,h+dh,w+dw,ci]   both MN-major, split-K
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

namespace {

enum { MODE_GEMM = 0, MODE_CONV = 1, MODE_WGRAD = 2 };
enum { DT_TF32 = 0, DT_BF16 = 1 };
enum { EPI_DIRECT = 0, EPI_STATS = 1, EPI_TMA = 2, EPI_TMA16 = 3 };
constexpr int kMaxTaps = 16;
constexpr int kStageABytes = 128 * 128;  // 128 rows x 128 B
constexpr int kStagingBytes = 8 * 4096;  // epilogue: one 32-row x 128-byte swizzled chunk per epilogue warp

struct GemmParams {
  int mode, dtype;
  int M, N;               // logical output extent (rows, cols)
  int block_n;            // 64 / 128 / 256
  int num_kb;             // k-blocks per tile
  int stages;
  int tiles_m, tiles_n, splits;
  int cluster;            // 1 = cta_group::1 (128 x block_n tile per CTA); 2 = cta_group::2: a CTA pair computes a
                          // 256 x block_n tile, each CTA stages its 128 rows of A and HALF of B in its own smem
  uint32_t stage_tx_bytes;  // bytes credited to the (leader's) full barrier per ring stage
  // ---- producer geometry
  int elems_per_128B;     // 32 (tf32) / 64 (bf16)
  int cblocks;            // CONV: channel blocks per tap (Cin / elems_per_128B)
  int ntaps;
  int tap_dh[kMaxTaps], tap_dw[kMaxTaps];
  int conv_stride;        // input coordinate = out * conv_stride + tap offset
  int tile_w, tile_h;     // output pixels per tile = tile_w * tile_h (=128 CONV, = kp WGRAD)
  int tiles_w, tiles_per_img;
  int Ho, Wo;             // output spatial extent (CONV/WGRAD pixel space)
  int kp;                 // WGRAD: pixels per k-block
  int wg_cin_blocks;      // WGRAD: N tiles per tap = Cin / block_n
  // ---- descriptors
  uint32_t idesc;
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;  // 16-byte units
  uint32_t a_kadv, b_kadv;              // 16-byte units per MMA k-step
  uint32_t layout_type;                 // UMMA LayoutType: 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  int mmas_per_kb;
  int a_boxes, b_boxes;                 // TMA boxes per stage
  uint32_t a_box_bytes, b_box_bytes;
  // ---- epilogue
  float* C;
  long ldc;
  const float* scale;     // per-column multiplier (optional)
  const float* bias;      // per-column addend (optional)
  const float* residual;  // same row mapping as C (optional)
  long ldr;
  int relu;
  int atomic;             // accumulate with red.global.add instead of store
  // row mapping for CONV: output pixel (n,oh,ow) -> row ((n*out_H + oh*out_s + out_oh)*out_W + ow*out_s + out_ow)
  int out_H, out_W, out_s, out_oh, out_ow;
  int out_bf16;           // store bf16 instead of fp32
  int stg_bufs;           // staging buffers per epilogue warp (1, or 2 for short-K launches whose epilogue is exposed)
  int epi_tma;            // 1: epilogue stages 32 x 32 chunks in shared memory and stores / reduces them by TMA (tma_c)
  // ---- tail split (K-major modes): work units [0, tail_first) are whole tiles; every tile >= tail_first (the
  // last, partial wave over the persistent grid) is cut into tail_s K-slices that run on different CTAs.  Slices
  // park their raw accumulators in tail_ws; the LAST slice to arrive (per epilogue warp, counted in tail_cnt) sums
  // all slices in index order (deterministic) and runs the normal epilogue.  tail_s = 0: off.
  int tail_first, tail_s;
  float4* tail_ws;
  int* tail_cnt;
  // ---- tail N-split (preferred when block_n = 256): every tile >= tail_first is cut into tail_p column pieces of
  // block_n / tail_p columns (own TMA box tma_bp, own instruction descriptor), one piece per CTA.  Same K order as a
  // whole tile -> bit-identical results, no fix-up.  tail_p = 0: off.  (tail_s and tail_p are mutually exclusive.)
  int tail_p;
  uint32_t idesc_piece;
  double* stats;          // optional [2*N]: += per-column sum and sum of squares of the stored values
                          // (train-mode BatchNorm statistics of the consumer, fused into the producer)
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// 2-SM (cta_group::2) loads: the box lands in THIS CTA's shared memory, the transaction bytes are credited to the
// LEADER CTA's mbarrier (peer bit of the shared::cluster address cleared), which is the one the MMA thread waits on.
__device__ __forceinline__ void tma_load_4d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                                int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// epilogue: shared -> global tile store / reduction through the TMA (bulk async-group completion)
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2,
                                                  int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read3() { asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// commit of cta_group::2 MMAs: one arrival on the barrier at the same offset in every CTA of cta_mask
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask)
               : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(rank)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
template <int DT>
__device__ __forceinline__ void umma_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  if (DT == 0) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
  }
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
template <int DT>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  if (DT == DT_TF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
  }
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_128B=2 [61,64)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo16, uint32_t sbo16, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)(lbo16 & 0x3fff) << 16;
  d |= (uint64_t)(sbo16 & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// one lane of a converged warp (the same one every time for a full mask)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

struct TileCoord {
  int tile_m, tile_n, split;
  int kb0, nkb;       // k-block range of this work unit
  int slice, tail;    // tail split: K-slice index and tail-tile index (slice = -1: whole tile)
  int col0, ncols;    // output columns of this work unit (K-major modes)
  int piece;          // 1 if this unit is a column piece of a tail tile
};

// work unit u of a cluster -> tile of this CTA (cluster of 2: consecutive M tiles, same N tile)
__device__ __forceinline__ TileCoord tile_coord(const GemmParams& p, int u, int cta_rank) {
  TileCoord c;
  c.kb0 = 0; c.nkb = p.num_kb; c.slice = -1; c.tail = 0;
  if (p.tail_s > 0 && u >= p.tail_first) {
    const int v = u - p.tail_first;
    c.tail = v / p.tail_s;
    c.slice = v - c.tail * p.tail_s;
    const int base = p.num_kb / p.tail_s, rem = p.num_kb - base * p.tail_s;
    c.kb0 = c.slice * base + min(c.slice, rem);
    c.nkb = base + (c.slice < rem ? 1 : 0);
    u = p.tail_first + c.tail;
  }
  int pc = 0;
  c.piece = 0;
  c.ncols = p.block_n;
  if (p.tail_p > 0 && u >= p.tail_first) {
    const int v = u - p.tail_first;
    const int tt = v / p.tail_p;
    pc = v - tt * p.tail_p;
    c.piece = 1;
    c.ncols = p.block_n / p.tail_p;
    u = p.tail_first + tt;
  }
  const int tiles_mp = (p.tiles_m + p.cluster - 1) / p.cluster;
  c.tile_n = u % p.tiles_n;
  const int r = u / p.tiles_n;
  c.tile_m = (r % tiles_mp) * p.cluster + cta_rank;
  c.split = r / tiles_mp;
  c.col0 = c.tile_n * p.block_n + pc * c.ncols;
  return c;
}

// Persistent: each CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...; the smem ring runs across tile
// boundaries and the accumulator is double-buffered in TMEM (2 x block_n columns), so the epilogue of tile i
// overlaps the TMA + MMA of tile i+1.
// CG = tcgen05 cta_group of this instantiation (a kernel may not mix cta_group::1 and ::2 instructions).
// EPI = epilogue variant (separate instantiations: 10 warps cap the kernel at 168 registers/thread, and sharing one
// body made the plain epilogue spill -- measured +3.8 ms/step): EPI_DIRECT = per-lane row stores (any alignment,
// bf16 output, ragged N), EPI_STATS = EPI_DIRECT + fused BatchNorm statistics, EPI_TMA = staged TMA store / reduce,
// EPI_TMA16 = the same for bf16 output (and bf16 residual): 32 x 32 chunks staged as 32 rows x 64 B with the 64-byte
// TMA swizzle.
// FEAT (TMA epilogues only): which optional epilogue stages are COMPILED IN -- bit 0 per-column scale / bias, bit 1
// residual, bit 2 ReLU, bit 3 fused BatchNorm statistics.  kFeatAll keeps every stage behind its runtime flag (the
// generic kernel); any other mask is a specialisation for launches that use exactly those stages, with the unused
// ones (and their predicated-off instructions: the short-K launches are bound by the ~400 issue slots per 32 x 32
// chunk of the generic epilogue) removed at compile time.
enum { FEAT_AFFINE = 1, FEAT_RES = 2, FEAT_RELU = 4, FEAT_STATS = 8, kFeatAll = 15 };

template <int DT, int CG, int EPI, int FEAT = kFeatAll>
__global__ void __launch_bounds__(320, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_bp,
               const __grid_constant__ GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x (A 16KB | B block_n*128B)] | epilogue staging 8 warps x 4 KB | barriers
  // 1024-byte alignment by pointer arithmetic on the __shared__ array (an integer round trip turns every later access
  // into a generic-address LD / ST: the SASS of the epilogue showed ST.E.128 / LD.E.128 instead of STS / LDS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t b_stage_bytes = (uint32_t)p.block_n * 128u;
  const uint32_t stage_bytes = kStageABytes + b_stage_bytes;
  uint8_t* staging = smem + (size_t)p.stages * stage_bytes;   // 1024-byte aligned (stage sizes are multiples of 1 KB)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + (size_t)kStagingBytes * p.stg_bufs);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full_bar = empty_bar + p.stages;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t acc_cols = p.block_n <= 32 ? 32u : (p.block_n <= 64 ? 64u : (p.block_n <= 128 ? 128u : 256u));
  const uint32_t tmem_cols = 2u * acc_cols;
  const int cta_rank = CG == 2 ? (int)cluster_ctarank() : 0;
  const int unit0 = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int unit_step = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int whole_tiles = ((p.tiles_m + CG - 1) / CG) * p.tiles_n * p.splits;
  const int tail_mult = p.tail_s > 0 ? p.tail_s : (p.tail_p > 0 ? p.tail_p : 1);
  const int total_tiles = p.tail_first + (whole_tiles - p.tail_first) * tail_mult;   // work units

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    if (EPI == EPI_TMA || EPI == EPI_TMA16) tma_prefetch_desc(&tma_c);
    if (p.tail_p > 0) tma_prefetch_desc(&tma_bp);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 8u * (uint32_t)CG);   // one arrival per epilogue warp (of both CTAs)
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (CG == 2) {
    __syncthreads();
    cluster_sync_all();   // both CTAs' barriers exist before any remote arrive / peer-credited TMA
    if (warp == 1) tmem_alloc_2sm(tmem_ptr_smem, tmem_cols);
  } else if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // =============================== TMA producer.  The whole warp walks the (warp-uniform) loops so that indices and
    // coordinates live in uniform registers; one elected lane issues the copies.
    {
      const bool leader = elect_one();
      const int E = p.elems_per_128B;
      int s = 0;          // ring slot / phase, carried across tiles
      uint32_t ph = 0;
      for (int t = unit0; t < total_tiles; t += unit_step) {
        const TileCoord tc = tile_coord(p, t, cta_rank);
        int n_img = 0, oh0 = 0, ow0 = 0;
        if (p.mode == MODE_CONV) {
          n_img = tc.tile_m / p.tiles_per_img;
          const int r = tc.tile_m - n_img * p.tiles_per_img;
          const int th = r / p.tiles_w;
          oh0 = th * p.tile_h;
          ow0 = (r - th * p.tiles_w) * p.tile_w;
        }
        int wg_tap = 0, wg_ci0 = 0;
        if (p.mode == MODE_WGRAD) {
          wg_tap = tc.tile_n / p.wg_cin_blocks;
          wg_ci0 = (tc.tile_n - wg_tap * p.wg_cin_blocks) * p.block_n;
        }
        for (int i = 0; i < tc.nkb; ++i) {
          mbar_wait(&empty_bar[s], ph ^ 1u);
          uint8_t* sa = smem + (size_t)s * stage_bytes;
          uint8_t* sb = sa + kStageABytes;
          const int kb = tc.split * p.num_kb + tc.kb0 + i;
          if (!leader) {
            // nothing to issue
          } else if (p.mode == MODE_GEMM) {
            if (cta_rank == 0)
              mbar_expect_tx(&full_bar[s], tc.piece ? (uint32_t)(kStageABytes + tc.ncols * 128) : p.stage_tx_bytes);
            if (CG == 2) {
              tma_load_4d_2sm(sa, &tma_a, &full_bar[s], kb * E, tc.tile_m * 128, 0, 0);
              tma_load_4d_2sm(sb, &tma_b, &full_bar[s], kb * E, tc.tile_n * p.block_n + cta_rank * (p.block_n >> 1), 0, 0);
            } else {
              tma_load_4d(sa, &tma_a, &full_bar[s], kb * E, tc.tile_m * 128, 0, 0);
              tma_load_4d(sb, tc.piece ? &tma_bp : &tma_b, &full_bar[s], kb * E, tc.col0, 0, 0);
            }
          } else if (p.mode == MODE_CONV) {
            if (cta_rank == 0)
              mbar_expect_tx(&full_bar[s], tc.piece ? (uint32_t)(kStageABytes + tc.ncols * 128) : p.stage_tx_bytes);
            const int tap = kb / p.cblocks, cb = kb - tap * p.cblocks;
            if (CG == 2) {
              tma_load_4d_2sm(sa, &tma_a, &full_bar[s], cb * E, ow0 * p.conv_stride + p.tap_dw[tap],
                              oh0 * p.conv_stride + p.tap_dh[tap], n_img);
              tma_load_4d_2sm(sb, &tma_b, &full_bar[s], kb * E, tc.tile_n * p.block_n + cta_rank * (p.block_n >> 1), 0, 0);
            } else {
              tma_load_4d(sa, &tma_a, &full_bar[s], cb * E, ow0 * p.conv_stride + p.tap_dw[tap],
                          oh0 * p.conv_stride + p.tap_dh[tap], n_img);
              tma_load_4d(sb, tc.piece ? &tma_bp : &tma_b, &full_bar[s], kb * E, tc.col0, 0, 0);
            }
          } else {
            // WGRAD: k-block = kp consecutive output pixels of one image row block
            if (cta_rank == 0) mbar_expect_tx(&full_bar[s], p.stage_tx_bytes);
            const int pix0 = kb * p.kp;
            const int img = pix0 / (p.Ho * p.Wo);
            const int rem = pix0 - img * (p.Ho * p.Wo);
            const int oh = rem / p.Wo, ow = rem - oh * p.Wo;
            if (CG == 2) {
              // each CTA: its 128 output channels of dY, and its half of the input-channel chunks of X
              for (int j = 0; j < p.a_boxes; ++j)
                tma_load_4d_2sm(sa + (size_t)j * p.a_box_bytes, &tma_a, &full_bar[s], tc.tile_m * 128 + j * E, pix0, 0, 0);
              const int hb = p.b_boxes >> 1;
              for (int j = 0; j < hb; ++j)
                tma_load_4d_2sm(sb + (size_t)j * p.b_box_bytes, &tma_b, &full_bar[s], wg_ci0 + (cta_rank * hb + j) * E,
                                ow * p.conv_stride + p.tap_dw[wg_tap], oh * p.conv_stride + p.tap_dh[wg_tap], img);
            } else {
              for (int j = 0; j < p.a_boxes; ++j)
                tma_load_4d(sa + (size_t)j * p.a_box_bytes, &tma_a, &full_bar[s], tc.tile_m * 128 + j * E, pix0, 0, 0);
              for (int j = 0; j < p.b_boxes; ++j)
                tma_load_4d(sb + (size_t)j * p.b_box_bytes, &tma_b, &full_bar[s], wg_ci0 + j * E,
                            ow * p.conv_stride + p.tap_dw[wg_tap], oh * p.conv_stride + p.tap_dh[wg_tap], img);
            }
          }
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (2-SM mode: the leader CTA issues for the pair).  All 32 lanes run the
    // loop -- descriptors and ring indices stay in uniform registers -- and one elected lane issues MMAs + commits
    // (with `if (lane == 0)` around the loop ptxas kept everything in vector registers and re-broadcast five
    // values per tcgen05.mma: the issue loop, not the tensor pipe, bounded the 128-wide tiles).
    if (cta_rank == 0) {
      const bool leader = elect_one();
      int s = 0;
      uint32_t ph = 0, lt = 0;
      for (int t = unit0; t < total_tiles; t += unit_step, ++lt) {
        const uint32_t acc = lt & 1u, acc_ph = (lt >> 1) & 1u;
        mbar_wait(&tmem_empty_bar[acc], acc_ph ^ 1u);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * acc_cols;
        const TileCoord mtc = tile_coord(p, t, 0);
        const int nkb = mtc.nkb;
        const uint32_t idesc = mtc.piece ? p.idesc_piece : p.idesc;
        for (int i = 0; i < nkb; ++i) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
          const uint32_t sb = sa + kStageABytes;
          const uint64_t adesc0 = make_smem_desc(sa, p.a_lbo, p.a_sbo, p.layout_type);
          const uint64_t bdesc0 = make_smem_desc(sb, p.b_lbo, p.b_sbo, p.layout_type);
          if (leader) {
            if (CG == 2) {
              for (int k = 0; k < p.mmas_per_kb; ++k)
                umma_2sm<DT>(tmem_d, adesc0 + (uint64_t)(k * p.a_kadv), bdesc0 + (uint64_t)(k * p.b_kadv), p.idesc,
                             (i | k) != 0 ? 1u : 0u);
              umma_commit_2sm(&empty_bar[s], 3);   // frees the stage in both CTAs' rings
            } else {
              for (int k = 0; k < p.mmas_per_kb; ++k)
                umma<DT>(tmem_d, adesc0 + (uint64_t)(k * p.a_kadv), bdesc0 + (uint64_t)(k * p.b_kadv), idesc,
                         (i | k) != 0 ? 1u : 0u);
              umma_commit(&empty_bar[s]);
            }
          }
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
        if (leader) {
          if (CG == 2) umma_commit_2sm(&tmem_full_bar[acc], 3);   // accumulator halves ready in both CTAs
          else umma_commit(&tmem_full_bar[acc]);
        }
      }
    }
  } else {
    // =============================== epilogue (warps 2..9): TMEM lane quarter q = warp & 3, column half h
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int m_local = q * 32 + lane;
    constexpr bool kGen = FEAT == kFeatAll;
    const bool f_scale = (FEAT & FEAT_AFFINE) && p.scale != nullptr;
    const bool f_bias = (FEAT & FEAT_AFFINE) && p.bias != nullptr;
    const bool f_res = (FEAT & FEAT_RES) && (!kGen || p.residual != nullptr);
    const bool f_relu = (FEAT & FEAT_RELU) && (!kGen || p.relu != 0);
    const bool f_stats = (FEAT & FEAT_STATS) && (!kGen || p.stats != nullptr);
    int sbuf = 0;   // staging buffer the next chunk uses (EPI_TMA)
    uint32_t lt = 0;
    // Fused BatchNorm statistics (EPI_TMA / EPI_TMA16): per-lane column sums are carried in registers across the tiles
    // of this CTA as long as the warp keeps working on the same columns (persistent CTAs revisit the same N tile
    // whenever gridDim is a multiple of tiles_n, which holds for every backbone shape but N = 2048), and flushed with
    // one pair of double atomics per column when the columns change or the CTA runs out of tiles: ~4 x fewer same-address
    // atomics than one flush per tile (640 -> 148 per column for M = 20480).
    float st_s0 = 0.f, st_s1 = 0.f, st_s2 = 0.f, st_s3 = 0.f, st_q0 = 0.f, st_q1 = 0.f, st_q2 = 0.f, st_q3 = 0.f;
    int st_base = -1;
    auto st_flush = [&]() {
      if (st_base >= 0) {
        const float ss[4] = {st_s0, st_s1, st_s2, st_s3}, qq[4] = {st_q0, st_q1, st_q2, st_q3};
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int col = st_base + ci * 32 + lane;
          if (col < p.N && (ss[ci] != 0.f || qq[ci] != 0.f)) {
            atomicAdd(p.stats + col, (double)ss[ci]);
            atomicAdd(p.stats + p.N + col, (double)qq[ci]);
          }
        }
      }
      st_s0 = st_s1 = st_s2 = st_s3 = st_q0 = st_q1 = st_q2 = st_q3 = 0.f;
    };
    auto st_add = [&](const int ci, const float sm, const float sq) {
      if (ci == 0) { st_s0 += sm; st_q0 += sq; }
      else if (ci == 1) { st_s1 += sm; st_q1 += sq; }
      else if (ci == 2) { st_s2 += sm; st_q2 += sq; }
      else { st_s3 += sm; st_q3 += sq; }
    };
    for (int t = unit0; t < total_tiles; t += unit_step, ++lt) {
      const TileCoord tc = tile_coord(p, t, cta_rank);
      const uint32_t acc = lt & 1u, acc_ph = (lt >> 1) & 1u;
      long row;
      bool row_ok;
      if (p.mode == MODE_CONV) {
        const int n_img = tc.tile_m / p.tiles_per_img;
        const int r = tc.tile_m - n_img * p.tiles_per_img;
        const int th = r / p.tiles_w;
        const int oh = th * p.tile_h + m_local / p.tile_w;
        const int ow = (r - th * p.tiles_w) * p.tile_w + m_local % p.tile_w;
        row = ((long)n_img * p.out_H + (long)oh * p.out_s + p.out_oh) * p.out_W + (long)ow * p.out_s + p.out_ow;
        row_ok = oh < p.Ho && ow < p.Wo && tc.tile_m < p.tiles_m;
      } else {
        row = (long)tc.tile_m * 128 + m_local;
        row_ok = row < p.M && tc.tile_m < p.tiles_m;
      }
      const int cols_per_warp = tc.ncols >> 1;   // unit width >= 64
      int col_base = tc.col0;
      if (p.mode == MODE_WGRAD) {
        const int tap = tc.tile_n / p.wg_cin_blocks;
        col_base = tap * (p.wg_cin_blocks * p.block_n) + (tc.tile_n - tap * p.wg_cin_blocks) * p.block_n;
      }
      col_base += half * cols_per_warp;
      if ((EPI == EPI_TMA || EPI == EPI_TMA16) && f_stats && col_base != st_base) {
        st_flush();
        st_base = col_base;
      }
      if (EPI == EPI_TMA16) {
        // ---------------- bf16 TMA-store epilogue (activations of the mixed-precision path).  Per 32-column chunk a
        // warp converts its 32 x 32 accumulator block to bf16 into a 2 KB staging buffer laid out as 32 rows x 64 B
        // with the 64-byte swizzle (16-byte chunk index ^ ((row >> 1) & 3)) and one lane stores it with a bulk tensor
        // copy.  The bf16 residual is read coalesced (8 rows x 64 B per instruction) into the same buffer one chunk
        // ahead.  BatchNorm statistics (optional) are taken from the ROUNDED values the consumer will read.
        uint8_t* const stg_base = staging + (size_t)(warp - 2) * p.stg_bufs * 4096;
        const int ch = lane & 3;
        int c_ow = 0, c_oh = 0, c_n = 0;
        int nvalid = 0;
        int rowi[4];
        const bool tile_ok = tc.tile_m < p.tiles_m;
        if (p.mode == MODE_CONV) {
          const int n_img = tc.tile_m / p.tiles_per_img;
          const int r = tc.tile_m - n_img * p.tiles_per_img;
          const int th = r / p.tiles_w;
          const int oh0 = th * p.tile_h, ow0 = (r - th * p.tiles_w) * p.tile_w;
          const int m0 = q * 32;
          c_n = n_img; c_oh = oh0 + m0 / p.tile_w; c_ow = ow0 + m0 % p.tile_w;
          nvalid = !tile_ok || c_oh >= p.Ho ? 0 : (p.tile_w >= 32 ? 32 : min(32, (p.Ho - c_oh) * p.tile_w));
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int m = m0 + i * 8 + (lane >> 2);
            const int oh = oh0 + m / p.tile_w, ow = ow0 + m % p.tile_w;
            rowi[i] = (oh < p.Ho && tile_ok)
                          ? ((n_img * p.out_H + oh * p.out_s + p.out_oh) * p.out_W + ow * p.out_s + p.out_ow) : -1;
          }
        } else {
          c_ow = tc.tile_m * 128 + q * 32;
          nvalid = tile_ok ? max(0, min(32, p.M - c_ow)) : 0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = c_ow + i * 8 + (lane >> 2);
            rowi[i] = (rr < p.M && tile_ok) ? rr : -1;
          }
        }
        const __nv_bfloat16* res16 = reinterpret_cast<const __nv_bfloat16*>(p.residual);
        const bool has_res = f_res;
        uint4 rn[4];
        if (has_res) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            rn[i] = (rowi[i] >= 0 && col_base < p.N)
                        ? __ldg(reinterpret_cast<const uint4*>(res16 + (long)rowi[i] * p.ldr + col_base) + ch)
                        : make_uint4(0u, 0u, 0u, 0u);
        }
        mbar_wait(&tmem_full_bar[acc], acc_ph);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + acc * acc_cols + (uint32_t)(half * cols_per_warp) + ((uint32_t)(q * 32) << 16);
        const int swr = (lane >> 1) & 3;
        for (int c = 0; c < cols_per_warp; c += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_acc + (uint32_t)c, v);
          if (c + 32 >= cols_per_warp) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CG == 2) mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
              else mbar_arrive(&tmem_empty_bar[acc]);
            }
          }
          const int n0 = col_base + c;
          if (n0 >= p.N || !tile_ok) continue;       // warp-uniform
          // a bf16 chunk needs 2 KB: every 4 KB fp32 staging slot holds two of them (2 or 4 bulk stores in flight)
          uint8_t* stg = stg_base + (size_t)sbuf * 2048;
          if (lane == 0) {
            if (p.stg_bufs == 2) bulk_wait_read3(); else bulk_wait_read1();
          }
          sbuf = (sbuf + 1) & (2 * p.stg_bufs - 1);
          __syncwarp();
          if (has_res) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = i * 8 + (lane >> 2);
              *reinterpret_cast<uint4*>(stg + r * 64 + ((ch ^ ((r >> 1) & 3)) << 4)) = rn[i];
            }
            if (c + 32 < cols_per_warp && n0 + 32 < p.N) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                rn[i] = rowi[i] >= 0
                            ? __ldg(reinterpret_cast<const uint4*>(res16 + (long)rowi[i] * p.ldr + n0 + 32) + ch)
                            : make_uint4(0u, 0u, 0u, 0u);
            }
            __syncwarp();
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[8 * j + e]);
            if (f_scale) {
              const float4 s0 = __ldg(reinterpret_cast<const float4*>(p.scale + n0) + 2 * j);
              const float4 s1 = __ldg(reinterpret_cast<const float4*>(p.scale + n0) + 2 * j + 1);
              f[0] *= s0.x; f[1] *= s0.y; f[2] *= s0.z; f[3] *= s0.w;
              f[4] *= s1.x; f[5] *= s1.y; f[6] *= s1.z; f[7] *= s1.w;
            }
            if (f_bias) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + 2 * j);
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + 2 * j + 1);
              f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
              f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
            }
            uint4* slot = reinterpret_cast<uint4*>(stg + lane * 64 + ((j ^ swr) << 4));
            if (has_res) {
              const uint4 r4 = *slot;
              const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 rf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rw[e]));
                f[2 * e] += rf.x;
                f[2 * e + 1] += rf.y;
              }
            }
            if (f_relu) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
            }
            uint32_t ow4[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
              ow4[e] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            *slot = make_uint4(ow4[0], ow4[1], ow4[2], ow4[3]);
          }
          fence_proxy_async();
          __syncwarp();
          if (f_stats) {
            // lane l sums column n0 + l of the staged (rounded) chunk over the rows that exist in the tensor
            float sm = 0.f, sq = 0.f;
            const uint8_t* colp = stg + ((lane & 7) << 1);
            const int jc = lane >> 3;
            if (nvalid == 32) {
              // full chunk: four independent accumulation chains (the single chain was ~130 dependent cycles per chunk)
              float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int r = 0; r < 32; ++r) {
                const float xv = __bfloat162float(
                    *reinterpret_cast<const __nv_bfloat16*>(colp + r * 64 + ((jc ^ ((r >> 1) & 3)) << 4)));
                s4[r & 3] += xv;
                q4[r & 3] = fmaf(xv, xv, q4[r & 3]);
              }
              sm = (s4[0] + s4[1]) + (s4[2] + s4[3]);
              sq = (q4[0] + q4[1]) + (q4[2] + q4[3]);
            } else {
#pragma unroll 8
              for (int r = 0; r < nvalid; ++r) {
                const float xv = __bfloat162float(
                    *reinterpret_cast<const __nv_bfloat16*>(colp + r * 64 + ((jc ^ ((r >> 1) & 3)) << 4)));
                sm += xv;
                sq = fmaf(xv, xv, sq);
              }
            }
            if (nvalid > 0) st_add(c >> 5, sm, sq);
          }
          if (lane == 0) {
            tma_store_4d(&tma_c, stg, n0, c_ow, c_oh, c_n);
            bulk_commit();
          }
        }
        continue;
      }
      if (EPI == EPI_TMA) {
        // ---------------- TMA-store epilogue.  Per 32-column chunk a warp transposes its 32 x 32 accumulator
        // block through a 128B-swizzled 4 KB staging buffer and ONE lane stores (or reduce-adds) it with a bulk
        // tensor copy: global writes are full 128-byte lines instead of 32 row-strided 16-byte pieces per
        // instruction, rows outside the tensor are clipped by the TMA.  The residual is read coalesced (4 rows x
        // 128 B per instruction) into the same buffer one chunk ahead.
        // (with two buffers per warp the store of chunk i is still reading its buffer while chunk i+1 is staged)
        uint8_t* const stg_base = staging + (size_t)(warp - 2) * p.stg_bufs * 4096;
        const int sw = lane & 7;
        int c_ow = 0, c_oh = 0, c_n = 0;        // TMA coordinates of this warp's 32 rows
        int nvalid = 0;                         // how many of them exist in the tensor (always a prefix)
        int rowi[8];                            // residual rows read by this lane (coalesced layout), -1 = none
        const bool tile_ok = tc.tile_m < p.tiles_m;
        if (p.mode == MODE_CONV) {
          const int n_img = tc.tile_m / p.tiles_per_img;
          const int r = tc.tile_m - n_img * p.tiles_per_img;
          const int th = r / p.tiles_w;
          const int oh0 = th * p.tile_h, ow0 = (r - th * p.tiles_w) * p.tile_w;
          const int m0 = q * 32;
          c_n = n_img; c_oh = oh0 + m0 / p.tile_w; c_ow = ow0 + m0 % p.tile_w;
          nvalid = !tile_ok || c_oh >= p.Ho ? 0 : (p.tile_w >= 32 ? 32 : min(32, (p.Ho - c_oh) * p.tile_w));
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int m = m0 + i * 4 + (lane >> 3);
            const int oh = oh0 + m / p.tile_w, ow = ow0 + m % p.tile_w;
            rowi[i] = (oh < p.Ho && tile_ok)
                          ? ((n_img * p.out_H + oh * p.out_s + p.out_oh) * p.out_W + ow * p.out_s + p.out_ow) : -1;
          }
        } else {
          c_ow = tc.tile_m * 128 + q * 32;
          nvalid = tile_ok ? max(0, min(32, p.M - c_ow)) : 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = c_ow + i * 4 + (lane >> 3);
            rowi[i] = (rr < p.M && tile_ok) ? rr : -1;
          }
        }
        const bool has_res = f_res;
        float4 rn[8];
        if (has_res) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            rn[i] = (rowi[i] >= 0 && col_base < p.N)
                        ? __ldg(reinterpret_cast<const float4*>(p.residual + (long)rowi[i] * p.ldr + col_base) + sw)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        mbar_wait(&tmem_full_bar[acc], acc_ph);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + acc * acc_cols + (uint32_t)(half * cols_per_warp) + ((uint32_t)(q * 32) << 16);
        const bool sliced = tc.slice >= 0;
        const int nchunks = cols_per_warp >> 5;
        // float4 index of (slice 0, this warp, chunk 0, j 0, lane) in the tail workspace; one slice = 32 * block_n float4
        const size_t ws_warp = ((size_t)tc.tail * p.tail_s * 8 + (size_t)(warp - 2)) * nchunks * 256 + lane;
        const size_t ws_slice = (size_t)8 * nchunks * 256;
        if (sliced) {
          // park this K-slice's raw accumulators (coalesced: instruction j writes 32 lanes x 16 B contiguous)
          float4* dst = p.tail_ws + ws_warp + (size_t)tc.slice * ws_slice;
          for (int c = 0; c < cols_per_warp; c += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_acc + (uint32_t)c, v);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              dst[(c >> 5) * 256 + j * 32] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                          __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
          }
          tc_fence_before();
          __threadfence();
          __syncwarp();
          int old = 0;
          if (lane == 0) {
            if (CG == 2) mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
            else mbar_arrive(&tmem_empty_bar[acc]);
            old = atomicAdd(p.tail_cnt + tc.tail * 8 + (warp - 2), 1);
          }
          old = __shfl_sync(0xffffffffu, old, 0);
          if (old != p.tail_s - 1) continue;          // another slice's warp finishes this part of the tile
          if (lane == 0) p.tail_cnt[tc.tail * 8 + (warp - 2)] = 0;   // self-cleaning for the next launch
          __threadfence();
        }
        for (int c = 0; c < cols_per_warp; c += 32) {
          uint32_t v[32];
          if (sliced) {
            // sum the slices in index order (deterministic); .cg loads: the data was written by other SMs
            const float4* src = p.tail_ws + ws_warp + (size_t)(c >> 5) * 256;
            float4 a4[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) a4[j] = __ldcg(src + j * 32);
            for (int sl = 1; sl < p.tail_s; ++sl) {
              const float4* s2 = src + (size_t)sl * ws_slice;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b4 = __ldcg(s2 + j * 32);
                a4[j].x += b4.x; a4[j].y += b4.y; a4[j].z += b4.z; a4[j].w += b4.w;
              }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              v[4 * j] = __float_as_uint(a4[j].x); v[4 * j + 1] = __float_as_uint(a4[j].y);
              v[4 * j + 2] = __float_as_uint(a4[j].z); v[4 * j + 3] = __float_as_uint(a4[j].w);
            }
          } else {
            tmem_ld32(tmem_acc + (uint32_t)c, v);
            if (c + 32 >= cols_per_warp) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if (CG == 2) mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
                else mbar_arrive(&tmem_empty_bar[acc]);
              }
            }
          }
          const int n0 = col_base + c;
          if (n0 >= p.N || !tile_ok) continue;       // warp-uniform
          // the bulk store that last used this staging buffer must have finished READING it
          uint8_t* stg = stg_base + (size_t)sbuf * 4096;
          if (lane == 0) {
            if (p.stg_bufs == 2) bulk_wait_read1(); else bulk_wait_read0();
          }
          if (p.stg_bufs == 2) sbuf ^= 1;
          __syncwarp();
          if (has_res) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = i * 4 + (lane >> 3);
              *reinterpret_cast<float4*>(stg + r * 128 + ((sw ^ (r & 7)) << 4)) = rn[i];
            }
            if (c + 32 < cols_per_warp && n0 + 32 < p.N) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                rn[i] = rowi[i] >= 0
                            ? __ldg(reinterpret_cast<const float4*>(p.residual + (long)rowi[i] * p.ldr + n0 + 32) + sw)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            __syncwarp();
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 f = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                   __uint_as_float(v[4 * j + 3]));
            if (f_scale) {
              const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.scale + n0) + j);
              f.x *= s4.x; f.y *= s4.y; f.z *= s4.z; f.w *= s4.w;
            }
            if (f_bias) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + j);
              f.x += b4.x; f.y += b4.y; f.z += b4.z; f.w += b4.w;
            }
            float4* slot = reinterpret_cast<float4*>(stg + lane * 128 + ((j ^ sw) << 4));
            if (has_res) {
              const float4 r4 = *slot;
              f.x += r4.x; f.y += r4.y; f.z += r4.z; f.w += r4.w;
            }
            if (f_relu) {
              f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); f.z = fmaxf(f.z, 0.f); f.w = fmaxf(f.w, 0.f);
            }
            *slot = f;
          }
          fence_proxy_async();
          __syncwarp();
          if (f_stats) {
            // fused BatchNorm statistics of the consumer: lane l sums column n0 + l of the staged chunk over the
            // rows that exist in the tensor (a prefix of the 32), conflict-free through the 128B swizzle
            float sm = 0.f, sq = 0.f;
            const uint8_t* colp = stg + ((lane & 3) << 2);
            const int jc = lane >> 2;
            if (nvalid == 32) {
              float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int r = 0; r < 32; ++r) {
                const float xv = *reinterpret_cast<const float*>(colp + r * 128 + ((jc ^ (r & 7)) << 4));
                s4[r & 3] += xv;
                q4[r & 3] = fmaf(xv, xv, q4[r & 3]);
              }
              sm = (s4[0] + s4[1]) + (s4[2] + s4[3]);
              sq = (q4[0] + q4[1]) + (q4[2] + q4[3]);
            } else {
#pragma unroll 8
              for (int r = 0; r < nvalid; ++r) {
                const float xv = *reinterpret_cast<const float*>(colp + r * 128 + ((jc ^ (r & 7)) << 4));
                sm += xv;
                sq = fmaf(xv, xv, sq);
              }
            }
            if (nvalid > 0) st_add(c >> 5, sm, sq);
          }
          if (lane == 0) {
            if (p.atomic) tma_reduce_add_4d(&tma_c, stg, n0, c_ow, c_oh, c_n);
            else tma_store_4d(&tma_c, stg, n0, c_ow, c_oh, c_n);
            bulk_commit();
          }
        }
        continue;
      }
      // bf16 output implies a bf16 residual (mixed-precision activations); fp32 otherwise
      const float* rbase = (p.residual && row_ok && !p.out_bf16) ? p.residual + row * p.ldr + col_base : nullptr;
      const __nv_bfloat16* rbase16 = (p.residual && row_ok && p.out_bf16)
                                         ? reinterpret_cast<const __nv_bfloat16*>(p.residual) + row * p.ldr + col_base
                                         : nullptr;
      const bool r_vec = rbase && ((reinterpret_cast<uintptr_t>(rbase) & 15) == 0) && (col_base + cols_per_warp <= p.N);
      // residual of the first chunk is requested before the accumulator is even ready
      float4 rn[8];
      if (r_vec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) rn[j] = __ldg(reinterpret_cast<const float4*>(rbase) + j);
      }
      mbar_wait(&tmem_full_bar[acc], acc_ph);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + acc * acc_cols + (uint32_t)(half * cols_per_warp) + ((uint32_t)(q * 32) << 16);
      for (int c = 0; c < cols_per_warp; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_acc + (uint32_t)c, v);
        float4 rc[8];
        if (r_vec) {
#pragma unroll
          for (int j = 0; j < 8; ++j) rc[j] = rn[j];
          if (c + 32 < cols_per_warp) {   // prefetch the next chunk's residual while this one is processed
#pragma unroll
            for (int j = 0; j < 8; ++j) rn[j] = __ldg(reinterpret_cast<const float4*>(rbase + c + 32) + j);
          }
        }
        if (c + 32 >= cols_per_warp) {
          // last chunk is in registers: hand the accumulator back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 2) mbar_arrive_cluster(&tmem_empty_bar[acc], 0);   // the leader's MMA thread waits on it
            else mbar_arrive(&tmem_empty_bar[acc]);
          }
        }
        const int n0 = col_base + c;
        if (n0 >= p.N) continue;                   // warp-uniform
        if (!row_ok && EPI != EPI_STATS) continue;         // with fused statistics every lane stays for the shuffles
        float* crow = p.C + row * p.ldc + n0;
        const bool full = (n0 + 32 <= p.N);
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (full) {
          if (p.scale) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.scale + n0 + j));
              f[j] *= s4.x; f[j + 1] *= s4.y; f[j + 2] *= s4.z; f[j + 3] *= s4.w;
            }
          }
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
              f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
            }
          }
          if (r_vec) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              f[4 * j] += rc[j].x; f[4 * j + 1] += rc[j].y; f[4 * j + 2] += rc[j].z; f[4 * j + 3] += rc[j].w;
            }
          } else if (rbase) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += __ldg(rbase + c + j);
          } else if (rbase16) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += __bfloat162float(rbase16[c + j]);
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (n0 + j < p.N) {
              float x = f[j];
              if (p.scale) x *= __ldg(p.scale + n0 + j);
              if (p.bias) x += __ldg(p.bias + n0 + j);
              if (rbase) x += __ldg(rbase + c + j);
              if (rbase16) x += __bfloat162float(rbase16[c + j]);
              if (p.relu) x = fmaxf(x, 0.0f);
              f[j] = x;
            }
          }
        }
        if (EPI == EPI_STATS) {
          // column sums over the 32 rows of this warp by recursive halving: after the loop lane l holds column l
          float sm[32], sq[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = row_ok ? f[j] : 0.0f;
            sm[j] = x;
            sq[j] = x * x;
          }
#pragma unroll
          for (int n = 16; n >= 1; n >>= 1) {
            const bool upper = (lane & n) != 0;
#pragma unroll
            for (int j = 0; j < n; ++j) {
              const float keep_s = upper ? sm[j + n] : sm[j], send_s = upper ? sm[j] : sm[j + n];
              const float keep_q = upper ? sq[j + n] : sq[j], send_q = upper ? sq[j] : sq[j + n];
              sm[j] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, n);
              sq[j] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, n);
            }
          }
          if (n0 + lane < p.N) {
            atomicAdd(p.stats + n0 + lane, (double)sm[0]);
            atomicAdd(p.stats + p.N + n0 + lane, (double)sq[0]);
          }
          if (!row_ok) continue;
        }
        if (p.out_bf16) {
          __nv_bfloat16* brow = reinterpret_cast<__nv_bfloat16*>(p.C) + row * p.ldc + n0;
          if (full && ((reinterpret_cast<uintptr_t>(brow) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint32_t w4[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[j + 2 * e], f[j + 2 * e + 1]);
                w4[e] = *reinterpret_cast<const uint32_t*>(&h2);
              }
              *reinterpret_cast<uint4*>(brow + j) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (full || n0 + j < p.N) brow[j] = __float2bfloat16(f[j]);
          }
        } else if (p.atomic) {
          if (full && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              atomicAdd(reinterpret_cast<float4*>(crow + j), make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (full || n0 + j < p.N) atomicAdd(crow + j, f[j]);
          }
        } else if (full && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(crow + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (full || n0 + j < p.N) crow[j] = f[j];
        }
      }
    }
    if ((EPI == EPI_TMA || EPI == EPI_TMA16) && f_stats) st_flush();   // this CTA has no more tiles
  }
  if ((EPI == EPI_TMA || EPI == EPI_TMA16) && warp >= 2 && lane == 0) bulk_wait_read0();   // staging buffers are read by in-flight bulk stores
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();   // nobody frees TMEM or exits while the peer still uses the pair
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    if (CG == 2) tmem_dealloc_2sm(tmem_base, tmem_cols);
    else tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 4-D tiled map; dims/strides innermost first; strides in bytes for dims 1..3.  atom32: 0 = SWIZZLE_128B,
// 1 = SWIZZLE_128B_ATOM_32B (MN-major fp32 operands), 2 = SWIZZLE_64B (bf16 epilogue staging)
int make_map(CUtensorMap* m, int dtype, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
             const uint32_t box[4], const uint32_t estr[4], int atom32 = 0) {
  EncodeTiledFn enc = get_encode();
  SN_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
  cuuint64_t gd[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t gs[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t es[4] = {estr[0], estr[1], estr[2], estr[3]};
  CUresult r = enc(m, dtype == DT_TF32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                   const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   atom32 == 2 ? CU_TENSOR_MAP_SWIZZLE_64B
                               : (atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SN_CHECK(r == CUDA_SUCCESS,
           "cuTensorMapEncodeTiled failed (%d): dims=%llu,%llu,%llu,%llu strides=%llu,%llu,%llu box=%u,%u,%u,%u es=%u,%u,%u,%u",
           (int)r, (unsigned long long)gd[0], (unsigned long long)gd[1], (unsigned long long)gd[2],
           (unsigned long long)gd[3], (unsigned long long)gs[0], (unsigned long long)gs[1], (unsigned long long)gs[2],
           bx[0], bx[1], bx[2], bx[3], es[0], es[1], es[2], es[3]);
  return 0;
}

uint32_t make_idesc(int dtype, int a_mn_major, int b_mn_major, int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                                   // c_format = F32
  const uint32_t fmt = dtype == DT_TF32 ? 2u : 1u;  // TF32 / BF16
  d |= fmt << 7;
  d |= fmt << 10;
  d |= (uint32_t)(a_mn_major ? 1 : 0) << 15;
  d |= (uint32_t)(b_mn_major ? 1 : 0) << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

int pick_stages(int block_n, int stg_bufs = 1) {
  // one persistent CTA per SM: fill shared memory with the operand ring
  const int stage = kStageABytes + block_n * 128;
  int s = (227 * 1024 - kStagingBytes * stg_bufs - 1024 - 256) / stage;
  return s > 8 ? 8 : s;
}

size_t smem_bytes(int stages, int block_n, int stg_bufs) {
  return (size_t)stages * (kStageABytes + block_n * 128) + (size_t)kStagingBytes * stg_bufs + 1024 + 256;
}

// Output tensor map for the TMA-store epilogue: fp32, 4-D (cols, w, h, n) with the row mapping's strides; box =
// 32 columns x the 32 rows one epilogue warp owns.  Returns 0 and sets p.epi_tma = 1 when the epilogue qualifies.
int make_out_map(CUtensorMap* mc, GemmParams& p) {
  const char* e = getenv("SNIPER_GEMM_TMA_STORE");   // read per call: the parity tests flip it between launches
  const bool enabled = !(e && e[0] == '0');
  memset(mc, 0, sizeof(*mc));
  p.epi_tma = 0;
  const int osz = p.out_bf16 ? 2 : 4;                 // bf16 output implies a bf16 residual
  const long ld_align = p.out_bf16 ? 8 : 4;             // 16-byte rows
  const bool ok = enabled && p.N % 32 == 0 && p.ldc % ld_align == 0 &&
                  (!p.out_bf16 || !p.atomic) &&
                  (reinterpret_cast<uintptr_t>(p.C) & 15) == 0 &&
                  (!p.residual || (p.ldr % ld_align == 0 && (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0)) &&
                  (!p.scale || (reinterpret_cast<uintptr_t>(p.scale) & 15) == 0) &&
                  (!p.bias || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
  if (!ok) return 0;
  const uint32_t ones[4] = {1, 1, 1, 1};
  const uint64_t rb = (uint64_t)p.ldc * osz;
  const int odt = p.out_bf16 ? DT_BF16 : DT_TF32;
  const int oswz = p.out_bf16 ? 2 : 0;
  if (p.mode == MODE_CONV) {
    const uint32_t bw = p.tile_w < 32 ? (uint32_t)p.tile_w : 32u;
    const uint64_t d[4] = {(uint64_t)p.N, (uint64_t)p.Wo, (uint64_t)p.Ho, (uint64_t)(p.M / (p.Ho * p.Wo))};
    const uint64_t st[3] = {rb * p.out_s, rb * p.out_s * p.out_W, rb * p.out_H * p.out_W};
    const uint32_t b[4] = {32, bw, 32u / bw, 1};
    const uint8_t* base = reinterpret_cast<const uint8_t*>(p.C) + ((long)p.out_oh * p.out_W + p.out_ow) * p.ldc * osz;
    if (make_map(mc, odt, base, d, st, b, ones, oswz)) return -1;
  } else {
    const uint64_t d[4] = {(uint64_t)p.N, (uint64_t)p.M, 1, 1};
    const uint64_t st[3] = {rb, rb * p.M, rb * p.M};
    const uint32_t b[4] = {32, 32, 1, 1};
    if (make_map(mc, odt, p.C, d, st, b, ones, oswz)) return -1;
  }
  p.epi_tma = 1;
  return 0;
}

// Tail-split workspaces: kTailSlots buffers per device (one per stream using the kernel concurrently) carved out of
// ONE caller-provided allocation (sniper_gemm_set_tail_workspace; the library itself never allocates).  One buffer
// holds the parked accumulators of at most 148 slices (128 x 256 fp32 each) + 148 x 8 arrival counters.  Without a
// registered workspace the K-slice tail split is simply not used (results differ only in summation order).
constexpr int kTailSlots = 4;
constexpr int kMaxDevices = 64;
constexpr size_t kTailWsBytes = (size_t)sn::kNumSMs * 128 * 256 * 4;
constexpr size_t kTailCntBytes = ((size_t)sn::kNumSMs * 8 * sizeof(int) + 255) / 256 * 256;
struct TailSlot {
  cudaStream_t stream;
  bool used;
  float4* ws;
  int* cnt;
};
struct TailDevice {
  bool ready;
  TailSlot slot[kTailSlots];
};
TailDevice g_tail[kMaxDevices];   // host-side table; the entry points are not re-entrant across host threads
                                  // (documented in include/sniper_b200.h): one host thread drives a device's launches

TailSlot* tail_slot(cudaStream_t stream) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices || !g_tail[dev].ready) return nullptr;
  TailSlot* sl = g_tail[dev].slot;
  for (int i = 0; i < kTailSlots; ++i)
    if (sl[i].used && sl[i].stream == stream) return &sl[i];
  for (int i = 0; i < kTailSlots; ++i)
    if (!sl[i].used) {
      sl[i].used = true;
      sl[i].stream = stream;
      return &sl[i];
    }
  return nullptr;   // more concurrent streams than slots: run without the tail split
}

// Decide the tail split of a K-major launch (see GemmParams::tail_*).  SNIPER_GEMM_TAIL=0 disables it.
void plan_tail(GemmParams& p, long tiles, long grid_units, cudaStream_t stream, bool dry = false) {
  p.tail_s = 0; p.tail_p = 0; p.tail_first = (int)tiles; p.tail_ws = nullptr; p.tail_cnt = nullptr;
  const char* e = getenv("SNIPER_GEMM_TAIL");   // 0 = off, 1 = K-slices only, 2 (default) = column pieces, else K-slices
  const int mode = e ? atoi(e) : 2;
  if (mode == 0) return;
  if (p.mode == MODE_WGRAD || p.cluster != 1 || !p.epi_tma || p.splits != 1) return;
  const long tail = tiles % grid_units;
  if (tail == 0 || tail * 2 > grid_units) return;
  if (mode == 2 && p.block_n == 256 && tiles > grid_units) {
    // column pieces: 4 x 64 columns if they fit one wave, else 2 x 128
    long pcs = tail * 4 <= grid_units ? 4 : 2;
    if (const char* m = getenv("SNIPER_GEMM_TAIL_MAXP")) pcs = pcs < atoi(m) ? pcs : atoi(m);
    if (pcs >= 2) {
      p.tail_p = (int)pcs;
      p.tail_first = (int)(tiles - tail);
      p.idesc_piece = make_idesc(p.dtype, 0, 0, 128, (int)(p.block_n / pcs));
      return;
    }
  }
  if (p.out_bf16) return;   // K-slices park fp32 accumulators and re-enter the fp32 epilogue: not wired for bf16 output
  // Cost model fitted to tools/gemm_time.py on B200: a 128 x 256 tile costs ~0.43 us per k-block; parking the
  // slices, the arrival counter and the late epilogue cost ~(9 + 2.4 S) us.  The split pays off for K >~ 2000 only
  // (K = 2304: 62 -> 56 us with S = 4; K = 1024: 41 -> 43 us, so it stays off there).
  const double t_tile = 0.43 * p.num_kb * p.block_n / 256.0;
  long max_s = grid_units / tail;
  if (max_s > p.num_kb / 4) max_s = p.num_kb / 4;      // at least 4 k-blocks per slice
  if (max_s > 8) max_s = 8;
  if (const char* m = getenv("SNIPER_GEMM_TAIL_MAXS")) max_s = max_s < atoi(m) ? max_s : atoi(m);
  long s = 0;
  double best = 3.0;                                   // minimum predicted gain, us
  for (long c = 2; c <= max_s; ++c) {
    const double gain = (1.0 - 1.0 / c) * t_tile - (9.0 + 2.4 * c);
    if (gain > best) { best = gain; s = c; }
  }
  if (s < 2) return;
  TailSlot* slot = dry ? nullptr : tail_slot(stream);
  if (!slot && !dry) return;
  p.tail_s = (int)s;
  p.tail_first = (int)(tiles - tail);
  p.tail_ws = slot ? slot->ws : nullptr;
  p.tail_cnt = slot ? slot->cnt : nullptr;
}

// K-major B operand [N rows, K] as launch() needs it to build the TMA map of a column piece
struct BOperand {
  const void* ptr;
  uint64_t K, N, ld_bytes;
};

int launch(const CUtensorMap& ma, const CUtensorMap& mb, GemmParams& p, dim3 tiles, cudaStream_t stream,
           const BOperand* bop = nullptr) {
  CUtensorMap mc, mbp;
  memset(&mbp, 0, sizeof(mbp));
  if (make_out_map(&mc, p)) return -1;
  p.tiles_m = (int)tiles.x; p.tiles_n = (int)tiles.y; p.splits = (int)tiles.z;
  long units = (long)((tiles.x + p.cluster - 1) / p.cluster) * tiles.y * tiles.z;
  const long max_units = sn::kNumSMs / p.cluster;
  plan_tail(p, units, max_units, stream);
  if (p.tail_p > 0 && !bop) { p.tail_p = 0; p.tail_first = (int)units; }
  if (p.tail_p > 0) {
    const uint64_t d[4] = {bop->K, bop->N, 1, 1};
    const uint64_t st[3] = {bop->ld_bytes, bop->ld_bytes * bop->N, bop->ld_bytes * bop->N};
    const uint32_t b[4] = {(uint32_t)p.elems_per_128B, (uint32_t)(p.block_n / p.tail_p), 1, 1};
    const uint32_t ones[4] = {1, 1, 1, 1};
    if (make_map(&mbp, p.dtype, bop->ptr, d, st, b, ones)) return -1;
  }
  if (p.tail_s > 0) units = p.tail_first + (units - p.tail_first) * p.tail_s;
  if (p.tail_p > 0) units = p.tail_first + (units - p.tail_first) * p.tail_p;
  dim3 grid((unsigned)((units < max_units ? units : max_units) * p.cluster), 1, 1);
  // Short-K launches are bound by the epilogue (a 128 x 256 tile's MMAs take 0.43 us per k-block, its stores ~7 us):
  // give each epilogue warp a second staging buffer so that staging chunk i+1 overlaps the bulk store of chunk i,
  // at the price of one ring stage.  SNIPER_GEMM_STG=1|2 forces a setting (A/B runs).
  p.stg_bufs = 1;
  if (p.mode != MODE_WGRAD && p.epi_tma) {
    p.stg_bufs = p.num_kb <= 24 ? 2 : 1;
    if (const char* e = getenv("SNIPER_GEMM_STG")) p.stg_bufs = atoi(e) == 2 ? 2 : 1;
    p.stages = pick_stages(p.block_n, p.stg_bufs);
  }
  const size_t smem = smem_bytes(p.stages, p.block_n, p.stg_bufs);
  typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const GemmParams);
  // [operand type][cluster - 1][epilogue]: the generic kernels (every optional epilogue stage behind its runtime flag)
  static const KernelFn kernels[2][2][4] = {
      {{gemm_tc_kernel<DT_TF32, 1, EPI_DIRECT>, gemm_tc_kernel<DT_TF32, 1, EPI_STATS>, gemm_tc_kernel<DT_TF32, 1, EPI_TMA>,
        gemm_tc_kernel<DT_TF32, 1, EPI_TMA16>},
       {gemm_tc_kernel<DT_TF32, 2, EPI_DIRECT>, gemm_tc_kernel<DT_TF32, 2, EPI_STATS>, gemm_tc_kernel<DT_TF32, 2, EPI_TMA>,
        gemm_tc_kernel<DT_TF32, 2, EPI_TMA16>}},
      {{gemm_tc_kernel<DT_BF16, 1, EPI_DIRECT>, gemm_tc_kernel<DT_BF16, 1, EPI_STATS>, gemm_tc_kernel<DT_BF16, 1, EPI_TMA>,
        gemm_tc_kernel<DT_BF16, 1, EPI_TMA16>},
       {gemm_tc_kernel<DT_BF16, 2, EPI_DIRECT>, gemm_tc_kernel<DT_BF16, 2, EPI_STATS>, gemm_tc_kernel<DT_BF16, 2, EPI_TMA>,
        gemm_tc_kernel<DT_BF16, 2, EPI_TMA16>}}};
  // [operand type][EPI_TMA | EPI_TMA16][mask]: 1-SM TMA-epilogue kernels specialised for the stage combinations the
  // training step launches most: plain (data gradients), statistics (conv -> train BN), residual (data gradient +
  // shortcut gradient), residual + statistics (conv3 + shortcut -> next unit's BN), scale/bias + ReLU (frozen BN folded)
  static const int kMasks[5] = {0, FEAT_STATS, FEAT_RES, FEAT_RES | FEAT_STATS, FEAT_AFFINE | FEAT_RELU};
#define SN_SPEC(DT, EPI)                                                                                         \
  {gemm_tc_kernel<DT, 1, EPI, 0>, gemm_tc_kernel<DT, 1, EPI, FEAT_STATS>, gemm_tc_kernel<DT, 1, EPI, FEAT_RES>,   \
   gemm_tc_kernel<DT, 1, EPI, FEAT_RES | FEAT_STATS>, gemm_tc_kernel<DT, 1, EPI, FEAT_AFFINE | FEAT_RELU>}
  static const KernelFn spec[2][2][5] = {{SN_SPEC(DT_TF32, EPI_TMA), SN_SPEC(DT_TF32, EPI_TMA16)},
                                         {SN_SPEC(DT_BF16, EPI_TMA), SN_SPEC(DT_BF16, EPI_TMA16)}};
#undef SN_SPEC
  // per-device one-time setup (a process may drive several GPUs)
  static bool attr_done[64] = {false};
  int dev = 0;
  SN_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_done[dev]) {
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) {
        for (int c = 0; c < 4; ++c)
          SN_CUDA(cudaFuncSetAttribute(kernels[a][b][c], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        for (int c = 0; c < 5; ++c)
          SN_CUDA(cudaFuncSetAttribute(spec[a][b][c], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      }
    attr_done[dev] = true;
  }
  const int ti = p.dtype == DT_TF32 ? 0 : 1;
  KernelFn fn = kernels[ti][p.cluster == 2 ? 1 : 0]
                       [p.epi_tma ? (p.out_bf16 ? EPI_TMA16 : EPI_TMA) : (p.stats ? EPI_STATS : EPI_DIRECT)];
  {
    const char* e = getenv("SNIPER_GEMM_SPEC");   // A/B: 0 = always the generic kernel
    if (p.epi_tma && p.cluster == 1 && !(e && e[0] == '0')) {
      const int mask = ((p.scale || p.bias) ? FEAT_AFFINE : 0) | (p.residual ? FEAT_RES : 0) | (p.relu ? FEAT_RELU : 0) |
                       (p.stats ? FEAT_STATS : 0);
      for (int c = 0; c < 5; ++c)
        if (kMasks[c] == mask) fn = spec[ti][p.out_bf16 ? 1 : 0][c];
    }
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = dim3(320, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)p.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  {
    const char* e = getenv("SNIPER_GEMM_CLUSTER_ATTR");   // A/B: force the (1,1,1) cluster attribute on 1-SM launches
    cfg.numAttrs = (p.cluster > 1 || (e && e[0] == '1')) ? 1 : 0;
  }
  SN_CUDA(cudaLaunchKernelEx(&cfg, fn, ma, mb, mc, mbp, p));
  SN_LAUNCH_CHECK();
  return 0;
}

// Tile width.  Measured on B200 (profiles/gemm_shapes_r01_v5_tilewidth.md): a 128 x 128 tile moves 1.33x the
// L2 -> shared-memory bytes per flop of a 128 x 256 tile and the K-major kernels are bound by exactly that traffic
// (conv 20480 x 512 x 27648: 674 TFLOP/s with block_n = 256, 276 with 128), so wide tiles win even where they
// quantise worse over the 148 SMs (M = 20480, N = 256: 160 tiles).  A wave-quantisation cost model
// (rounds x (bn + overhead)) was tried and lost 8 ms/step.  SNIPER_GEMM_BN=<64|128|256> forces a width (A/B runs).
int pick_block_n(int N, long tiles_m) {
  (void)tiles_m;
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("SNIPER_GEMM_BN");
    forced = e ? atoi(e) : 0;
  }
  const int cap = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
  if (forced == 64 || forced == 128 || forced == 256) return forced < cap ? forced : cap;
  if (N <= 128) return cap;
  if (N % 256 == 0 || N > 1024) return 256;
  return 128;
}

// cta_group::2: a CTA pair (cluster of 2, consecutive M tiles) issues 256 x block_n MMAs from the leader CTA; each
// CTA stages its own 128 rows of A and HALF of B.  A 128 x 256 fp32-operand tile on ONE SM needs 96 B/clk of TMA fill
// plus 96 B/clk of MMA operand reads from a 128 B/clk shared-memory port (<= 67 % tensor utilisation); the pair needs
// 64 + 64.  (A plain cluster with TMA multicast of B was measured ~20 % slower, profiles/gemm_shapes_r01_v3_multicast.md:
// it saves L2 reads, not shared-memory traffic.)  Measured (profiles/gemm_shapes_r01_v4_2sm.md): correct, but the
// K-major shapes run 10-15 % slower than the 1-SM kernel (the largest conv already reaches 674 TFLOP/s = 79 % of
// the measured TF32 burst peak with cta_group::1, i.e. it is not operand-bound), so it is opt-in: SNIPER_GEMM_2SM=1.
int pick_cluster(int tiles_m) {
  const char* e = getenv("SNIPER_GEMM_2SM");   // read per call (A/B runs flip it between launches)
  const int enabled = (e && e[0] == '1') ? 1 : 0;
  return (enabled && tiles_m >= 2) ? 2 : 1;
}

void fill_kmajor(GemmParams& p, int dtype, int block_n) {
  p.dtype = dtype;
  p.block_n = block_n;
  p.elems_per_128B = dtype == DT_TF32 ? 32 : 64;
  p.idesc = make_idesc(dtype, 0, 0, 128 * p.cluster, block_n);
  p.a_lbo = 1; p.a_sbo = 64; p.b_lbo = 1; p.b_sbo = 64;  // SBO = 8 rows x 128 B
  p.a_kadv = 2; p.b_kadv = 2;                            // 32 B per UMMA_K step
  p.layout_type = 2;
  p.mmas_per_kb = 4;
  p.a_boxes = 1; p.b_boxes = 1;
  p.a_box_bytes = kStageABytes;
  p.b_box_bytes = (uint32_t)(block_n / p.cluster) * 128u;   // 2-SM: each CTA stages half of the B rows
  p.stage_tx_bytes = (uint32_t)p.cluster * (kStageABytes + p.b_box_bytes);
  p.stages = pick_stages(block_n);
}

struct Epilogue {
  const float* scale;
  const float* bias;
  const float* residual;
  long ldr;
  int relu;
  int atomic;
  int out_bf16;
};

}  // namespace

extern "C" {

// C[M,N] (ldc) = epilogue(A[M,K] (lda) * B[N,K]^T (ldb)).  dtype 0: fp32 storage / TF32 math, 1: bf16.
// K must be a multiple of 32 (tf32) / 64 (bf16) elements; lda/ldb in elements, 16-byte aligned rows.
int sniper_gemm_nt(const void* A, long lda, const void* B, long ldb, float* C, long ldc, int M, int N, int K,
                   int dtype, const float* scale, const float* bias, const float* residual, long ldr, int relu,
                   int accumulate, int out_bf16, double* stats, void* stream) {
  SN_CHECK(dtype == DT_TF32 || dtype == DT_BF16, "gemm: dtype must be 0 (tf32) or 1 (bf16)");
  const int esz = dtype == DT_TF32 ? 4 : 2;
  const int E = 128 / esz;
  SN_CHECK(K % E == 0 && K > 0, "gemm: K (%d) must be a positive multiple of %d", K, E);
  SN_CHECK(M > 0 && N > 0, "gemm: empty problem");
  SN_CHECK((lda * esz) % 16 == 0 && (ldb * esz) % 16 == 0, "gemm: row strides must be 16-byte multiples");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  const int bn = pick_block_n(N, sn::div_up(M, 128));
  p.cluster = pick_cluster(sn::div_up(M, 128));
  fill_kmajor(p, dtype, bn);
  p.mode = MODE_GEMM; p.M = M; p.N = N; p.num_kb = K / E;
  p.C = C; p.ldc = ldc; p.scale = scale; p.bias = bias; p.residual = residual; p.ldr = ldr; p.relu = relu;
  p.atomic = accumulate; p.out_bf16 = out_bf16; p.stats = stats;
  CUtensorMap ma, mb;
  const uint32_t ones[4] = {1, 1, 1, 1};
  {
    const uint64_t d[4] = {(uint64_t)K, (uint64_t)M, 1, 1};
    const uint64_t s[3] = {(uint64_t)lda * esz, (uint64_t)lda * esz * M, (uint64_t)lda * esz * M};
    const uint32_t b[4] = {(uint32_t)E, 128, 1, 1};
    if (make_map(&ma, dtype, A, d, s, b, ones)) return -1;
  }
  {
    const uint64_t d[4] = {(uint64_t)K, (uint64_t)N, 1, 1};
    const uint64_t s[3] = {(uint64_t)ldb * esz, (uint64_t)ldb * esz * N, (uint64_t)ldb * esz * N};
    const uint32_t b[4] = {(uint32_t)E, (uint32_t)(bn / p.cluster), 1, 1};
    if (make_map(&mb, dtype, B, d, s, b, ones)) return -1;
  }
  dim3 grid(sn::div_up(M, 128), sn::div_up(N, bn), 1);
  const BOperand bop = {B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * esz};
  return launch(ma, mb, p, grid, (cudaStream_t)stream, &bop);
}

// NHWC implicit-GEMM convolution forward (also used for stride-1 data gradients with flipped weights):
//   Y[n,oh,ow,co] = epi( sum_{t<ntaps, c<Cin} X[n, oh*stride + dh[t], ow*stride + dw[t], c] * Wt[co, t*Cin + c] )
// X: [NB,H,W,Cin] contiguous; Wt: [Cout, ntaps*Cin]; Y row (n,oh,ow) is written at
// ((n*out_H + oh*out_s + out_oh)*out_W + ow*out_s + out_ow) * ldc.  Out-of-range taps read zeros (TMA fill).
int sniper_conv2d_nhwc(const void* X, long x_ld, int NB, int H, int W, int Cin, const void* Wt, int Cout, int ntaps,
                       const int* tap_dh, const int* tap_dw, int stride, int Ho, int Wo, float* Y, long ldc,
                       int out_H, int out_W, int out_s, int out_oh, int out_ow, int dtype, const float* scale,
                       const float* bias, const float* residual, long ldr, int relu, int accumulate, int out_bf16,
                       double* stats, void* stream) {
  SN_CHECK(dtype == DT_TF32 || dtype == DT_BF16, "conv: dtype must be 0 (tf32) or 1 (bf16)");
  const int esz = dtype == DT_TF32 ? 4 : 2;
  const int E = 128 / esz;
  SN_CHECK(Cin % E == 0, "conv: Cin (%d) must be a multiple of %d", Cin, E);
  SN_CHECK(ntaps >= 1 && ntaps <= kMaxTaps, "conv: ntaps (%d) out of range", ntaps);
  SN_CHECK(stride == 1 || stride == 2, "conv: stride must be 1 or 2");
  // 128 output pixels per tile as tile_h rows x tile_w columns, tile_w = the largest power of two <= 128 dividing Wo
  // (training shapes: Wo = 32 ... 128 -> whole rows; inference canvases of any width that is a multiple of 8 pixels at
  // this layer's stride: narrower, taller tiles; rows past Ho are clipped by the TMA)
  int tile_w = 128;
  while (tile_w > 1 && Wo % tile_w != 0) tile_w >>= 1;
  SN_CHECK(tile_w >= 8, "conv: Wo (%d) must be a multiple of 8", Wo);
  int tile_h = 128 / tile_w;
  SN_CHECK(tile_w * stride <= 256, "conv: TMA box too wide");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  const int bn = pick_block_n(Cout, (long)NB * (Wo / tile_w) * sn::div_up(Ho, tile_h));
  p.cluster = pick_cluster(NB * (Wo / tile_w) * sn::div_up(Ho, tile_h));
  fill_kmajor(p, dtype, bn);
  p.mode = MODE_CONV; p.N = Cout; p.cblocks = Cin / E; p.ntaps = ntaps; p.num_kb = ntaps * p.cblocks;
  for (int t = 0; t < ntaps; ++t) { p.tap_dh[t] = tap_dh[t]; p.tap_dw[t] = tap_dw[t]; }
  p.conv_stride = stride; p.tile_w = tile_w; p.tile_h = tile_h;
  p.tiles_w = Wo / tile_w; p.tiles_per_img = p.tiles_w * sn::div_up(Ho, tile_h);
  p.Ho = Ho; p.Wo = Wo; p.M = NB * Ho * Wo;
  p.C = Y; p.ldc = ldc; p.scale = scale; p.bias = bias; p.residual = residual; p.ldr = ldr; p.relu = relu;
  p.atomic = accumulate; p.out_bf16 = out_bf16; p.stats = stats;
  p.out_H = out_H; p.out_W = out_W; p.out_s = out_s; p.out_oh = out_oh; p.out_ow = out_ow;
  CUtensorMap ma, mb;
  {
    SN_CHECK(x_ld >= Cin && (x_ld * esz) % 16 == 0, "conv: bad x_ld");
    const uint64_t d[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)NB};
    const uint64_t s[3] = {(uint64_t)x_ld * esz, (uint64_t)x_ld * esz * W, (uint64_t)x_ld * esz * W * H};
    // with elementStrides the box extent is in un-strided elements
    const uint32_t b[4] = {(uint32_t)E, (uint32_t)(tile_w * stride), (uint32_t)(tile_h * stride), 1};
    const uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
    if (make_map(&ma, dtype, X, d, s, b, es)) return -1;
  }
  {
    const uint64_t K = (uint64_t)ntaps * Cin;
    const uint64_t d[4] = {K, (uint64_t)Cout, 1, 1};
    const uint64_t s[3] = {K * esz, K * esz * Cout, K * esz * Cout};
    const uint32_t b[4] = {(uint32_t)E, (uint32_t)(bn / p.cluster), 1, 1};
    const uint32_t ones[4] = {1, 1, 1, 1};
    if (make_map(&mb, dtype, Wt, d, s, b, ones)) return -1;
  }
  dim3 grid(NB * p.tiles_per_img, sn::div_up(Cout, bn), 1);
  const BOperand bop = {Wt, (uint64_t)ntaps * Cin, (uint64_t)Cout, (uint64_t)ntaps * Cin * esz};
  return launch(ma, mb, p, grid, (cudaStream_t)stream, &bop);
}

// Weight gradient:  dW[co, t*Cin + ci] += sum_{n,oh,ow} dY[(n,oh,ow), co] * X[n, oh*stride+dh[t], ow*stride+dw[t], ci]
// dY: [NB*Ho*Wo, Cout] contiguous rows (ld = Cout); X: [NB,H,W,Cin]; dW must be zero-initialised by the caller
// (accumulated with red.global.add across split-K CTAs).  ntaps = 1, dh=dw=0, H=Ho, W=Wo gives dY^T * X (FC layers).
int sniper_conv2d_wgrad_nhwc(const void* dY, long dy_ld, const void* X, long x_ld, int NB, int H, int W, int Cin,
                             int Cout, int ntaps,
                             const int* tap_dh, const int* tap_dw, int stride, int Ho, int Wo, float* dW, int dtype,
                             int splits, void* stream) {
  SN_CHECK(dtype == DT_TF32 || dtype == DT_BF16, "wgrad: dtype must be 0 (tf32) or 1 (bf16)");
  const int esz = dtype == DT_TF32 ? 4 : 2;
  const int E = 128 / esz;       // elements per 128-byte MN chunk
  const int kp = dtype == DT_TF32 ? 32 : 64;  // pixels per k-block
  SN_CHECK(Cin % 64 == 0, "wgrad: Cin (%d) must be a multiple of 64", Cin);
  SN_CHECK(Cout % E == 0, "wgrad: Cout (%d) must be a multiple of %d", Cout, E);
  SN_CHECK(ntaps >= 1 && ntaps <= kMaxTaps, "wgrad: ntaps out of range");
  const long pixels = (long)NB * Ho * Wo;
  // k-blocks are kp consecutive output pixels.  Either they tile rows exactly, or the problem is a single
  // row (NB == 1, Ho == 1: FC layers viewed as [1,1,N,C]) where the tail is zero-filled by TMA.
  const bool single_row = (NB == 1 && Ho == 1);
  SN_CHECK(single_row || ((Wo % kp == 0 || kp % Wo == 0) && pixels % kp == 0 && (Wo >= kp || (Ho * Wo) % kp == 0)),
           "wgrad: output %dx%d does not tile by %d pixels", Ho, Wo, kp);
  const int bw = Wo >= kp ? kp : Wo, bh = kp / bw;
  // wide N tiles: fp32 operands make the 128x128 tile L2-bandwidth bound (128 B/clk/SM of operand traffic);
  // 128x256 needs 96 B/clk and halves the number of split-K atomics
  const int bn = Cin % 256 == 0 ? 256 : (Cin % 128 == 0 ? 128 : 64);
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.mode = MODE_WGRAD; p.dtype = dtype; p.block_n = bn; p.elems_per_128B = E;
  p.M = Cout; p.N = ntaps * Cin;
  p.idesc = make_idesc(dtype, 1, 1, 128, bn);
  const uint32_t chunk_bytes = (uint32_t)kp * 128u;
  // MN-major: LBO = stride between 128-byte MN atoms, SBO = stride between k-row groups.  32-bit
  // operands must use the 128B swizzle with 32-byte atoms (UMMA SWIZZLE_128B_BASE32B, TMA
  // SWIZZLE_128B_ATOM_32B): k-row groups are 4 rows (512 B) instead of 8 (1024 B).
  const int atom32 = dtype == DT_TF32 ? 1 : 0;
  p.layout_type = atom32 ? 1u : 2u;
  const uint32_t sbo = atom32 ? 32u : 64u;
  p.a_lbo = chunk_bytes >> 4; p.a_sbo = sbo; p.b_lbo = chunk_bytes >> 4; p.b_sbo = sbo;
  const int umma_k = 32 / esz;                       // 8 (tf32) / 16 (bf16) k-rows per MMA
  p.a_kadv = (uint32_t)(umma_k * 128) >> 4; p.b_kadv = p.a_kadv;
  p.mmas_per_kb = kp / umma_k;
  p.a_boxes = 128 / E; p.b_boxes = bn / E;
  p.a_box_bytes = chunk_bytes; p.b_box_bytes = chunk_bytes;
  p.cluster = (p.b_boxes % 2 == 0) ? pick_cluster(Cout / 128 + (Cout % 128 ? 1 : 0)) : 1;
  p.stage_tx_bytes = (uint32_t)p.cluster * (uint32_t)(p.a_boxes + p.b_boxes / p.cluster) * chunk_bytes;
  p.idesc = make_idesc(dtype, 1, 1, 128 * p.cluster, bn);
  p.stages = pick_stages(bn);
  p.ntaps = ntaps;
  for (int t = 0; t < ntaps; ++t) { p.tap_dh[t] = tap_dh[t]; p.tap_dw[t] = tap_dw[t]; }
  p.conv_stride = stride; p.Ho = Ho; p.Wo = Wo; p.kp = kp; p.wg_cin_blocks = Cin / bn;
  const long total_kb = (pixels + kp - 1) / kp;
  if (splits < 1) {
    // auto: fill one wave of 148 persistent CTAs (two if the tiles are tiny), at least 8 k-blocks per split
    const long base_tiles = (long)(Cout / 128 + (Cout % 128 ? 1 : 0)) * ntaps * p.wg_cin_blocks;
    long s = base_tiles >= sn::kNumSMs ? 1 : (sn::kNumSMs + base_tiles / 2) / base_tiles;
    if (s > total_kb / 8) s = total_kb / 8;
    splits = s < 1 ? 1 : (int)s;
  }
  while (splits > 1 && total_kb % splits != 0) --splits;
  p.num_kb = (int)(total_kb / splits);
  p.C = dW; p.ldc = (long)ntaps * Cin; p.atomic = 1;
  CUtensorMap ma, mb;
  {
    SN_CHECK(dy_ld >= Cout && x_ld >= Cin && (dy_ld * esz) % 16 == 0 && (x_ld * esz) % 16 == 0, "wgrad: bad ld");
    const uint64_t d[4] = {(uint64_t)Cout, (uint64_t)pixels, 1, 1};
    const uint64_t s[3] = {(uint64_t)dy_ld * esz, (uint64_t)dy_ld * esz * pixels, (uint64_t)dy_ld * esz * pixels};
    const uint32_t b[4] = {(uint32_t)E, (uint32_t)kp, 1, 1};
    const uint32_t ones[4] = {1, 1, 1, 1};
    if (make_map(&ma, dtype, dY, d, s, b, ones, atom32)) return -1;
  }
  {
    const uint64_t d[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)NB};
    const uint64_t s[3] = {(uint64_t)x_ld * esz, (uint64_t)x_ld * esz * W, (uint64_t)x_ld * esz * W * H};
    const uint32_t b[4] = {(uint32_t)E, (uint32_t)(bw * stride), (uint32_t)(bh * stride), 1};
    const uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
    if (make_map(&mb, dtype, X, d, s, b, es, atom32)) return -1;
  }
  dim3 grid(Cout / 128 + (Cout % 128 ? 1 : 0), ntaps * p.wg_cin_blocks, splits);
  return launch(ma, mb, p, grid, (cudaStream_t)stream);
}

// Bytes sniper_gemm_set_tail_workspace expects (per device).
size_t sniper_gemm_tail_workspace_bytes(void) { return (size_t)kTailSlots * (kTailWsBytes + kTailCntBytes); }

// Registers a caller-owned, ZERO-FILLED device buffer of sniper_gemm_tail_workspace_bytes() bytes for `device`; it must
// outlive every later launch on that device.  Enables the K-slice split of the last partial wave (GemmParams::tail_s).
// ws = NULL unregisters.
int sniper_gemm_set_tail_workspace(void* ws, size_t bytes, int device) {
  SN_CHECK(device >= 0 && device < kMaxDevices, "gemm_set_tail_workspace: device %d out of range", device);
  TailDevice& d = g_tail[device];
  d.ready = false;
  if (ws == nullptr) return 0;
  SN_CHECK(bytes >= sniper_gemm_tail_workspace_bytes(), "gemm_set_tail_workspace: %zu bytes given, %zu needed", bytes,
           sniper_gemm_tail_workspace_bytes());
  SN_CHECK((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "gemm_set_tail_workspace: buffer must be 256-byte aligned");
  uint8_t* p8 = static_cast<uint8_t*>(ws);
  for (int i = 0; i < kTailSlots; ++i) {
    d.slot[i].stream = nullptr;
    d.slot[i].used = false;
    d.slot[i].ws = reinterpret_cast<float4*>(p8 + (size_t)i * (kTailWsBytes + kTailCntBytes));
    d.slot[i].cnt = reinterpret_cast<int*>(p8 + (size_t)i * (kTailWsBytes + kTailCntBytes) + kTailWsBytes);
  }
  d.ready = true;
  return 0;
}

// Host-only query of the launch plan sniper_gemm_nt would use for C[M,N] = A[M,K] * B[N,K]^T with 16-byte aligned,
// contiguous operands (no GPU needed; exercised by the CPU tests).  out[8] = {block_n, ring stages, staging buffers per
// epilogue warp, tiles, persistent grid, tail mode (0 none, 1 K-slices, 2 column pieces), first tail tile, slices or
// pieces per tail tile}.
int sniper_gemm_plan(int M, int N, int K, int dtype, int* out) {
  SN_CHECK(dtype == DT_TF32 || dtype == DT_BF16, "gemm_plan: dtype must be 0 (tf32) or 1 (bf16)");
  const int E = dtype == DT_TF32 ? 32 : 64;
  SN_CHECK(M > 0 && N > 0 && K > 0 && K % E == 0, "gemm_plan: K (%d) must be a positive multiple of %d", K, E);
  GemmParams p;
  memset(&p, 0, sizeof(p));
  const int bn = pick_block_n(N, sn::div_up(M, 128));
  p.cluster = pick_cluster(sn::div_up(M, 128));
  fill_kmajor(p, dtype, bn);
  p.mode = MODE_GEMM; p.M = M; p.N = N; p.num_kb = K / E;
  p.epi_tma = (N % 32 == 0) ? 1 : 0;
  p.splits = 1;
  const long tiles = (long)sn::div_up(sn::div_up(M, 128), p.cluster) * sn::div_up(N, bn);
  const long max_units = sn::kNumSMs / p.cluster;
  plan_tail(p, tiles, max_units, nullptr, /*dry=*/true);
  long units = tiles;
  int mode = 0, factor = 1;
  if (p.tail_s > 0) { mode = 1; factor = p.tail_s; }
  if (p.tail_p > 0) { mode = 2; factor = p.tail_p; }
  units = p.tail_first + (tiles - p.tail_first) * factor;
  int stg = 1;
  if (p.epi_tma) stg = p.num_kb <= 24 ? 2 : 1;
  out[0] = bn;
  out[1] = pick_stages(bn, stg);
  out[2] = stg;
  out[3] = (int)tiles;
  out[4] = (int)((units < max_units ? units : max_units) * p.cluster);
  out[5] = mode;
  out[6] = p.tail_first;
  out[7] = factor;
  return 0;
}

}  // extern "C"
