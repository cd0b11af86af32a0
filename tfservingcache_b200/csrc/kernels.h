// Device kernels of the executor. All fp32, sm_100a.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace tfsc {

constexpr int kMaxRowsPerLaunch = 8;  // rows handled by one streaming pass over W (SIMT path)

// X1: y = a*x + b (a, b device scalars). half_plus_two (deploy/docker-compose/readme.md:40-42).
cudaError_t launch_affine(const float* x, float* y, int64_t n, const float* a, const float* b, cudaStream_t s);

// X2: y[rows,n] = act(x[rows,k] W[k,n] + bias[n]); W row-major [k,n]. Streams W exactly once per
// group of <= kMaxRowsPerLaunch rows. workspace: dense_workspace_bytes(rows,k,n), zero-initialised
// counters are maintained by the kernel itself (self-resetting).
size_t dense_workspace_bytes(int rows, int k, int n);
cudaError_t launch_dense(const float* x, const float* w, const float* bias, float* y, int rows, int k, int n,
                         bool relu, void* workspace, size_t workspace_bytes, cudaStream_t s, int variant = 0);
// variant: 0 auto (TFSC_DENSE_VARIANT, default LDG stream for <= 8 rows + tensor cores above), 1 LDG stream only,
// 2 / 4 bulk-copy (TMA) ring for the <= 8-row passes (tensor cores above, as in auto), 3 tensor cores for every row count,
// 5 cluster-pair kernel for the <= 8-row passes (experimental)

// X3: tcgen05/TMEM 3xTF32 path for 9..64 rows per pass (dense_tc.cu)
bool dense_tc_supported(int rows, int k, int n, const float* w, const float* x, const float* bias, const float* y);
size_t dense_tc_workspace_bytes(int k, int n);
cudaError_t launch_dense_tc(const float* x, const float* w, const float* bias, float* y, int rows, int k, int n,
                            bool relu, void* workspace, size_t workspace_bytes, cudaStream_t s);

// X2, cluster-pair variant (dense_cluster.cu, EXPERIMENTAL, variant 5): two CTAs of a cluster split K and meet in
// distributed shared memory -- no split-K workspace, no atomics. rows <= 8, n % 4 == 0, k % 4 == 0, k >= 128.
bool dense_cluster_supported(int rows, int k, int n, const float* w, const float* x, const float* bias, const float* y);
cudaError_t launch_dense_cluster(const float* x, const float* w, const float* bias, float* y, int rows, int k, int n, bool relu,
                                 cudaStream_t s);

// X4/X5 building blocks (nn_kernels.cu): act 0 none / 1 relu / 2 gelu(erf)
cudaError_t launch_gemm(const float* A, const float* B, const float* bias, const float* R, float* C, int M, int N, int K,
                        int lda, int act, cudaStream_t s);
cudaError_t launch_im2col(const float* x, float* col, int Bn, int H, int W, int C, int KH, int KW, int stride, int pad,
                          int OH, int OW, int ldc, cudaStream_t s);
cudaError_t launch_maxpool(const float* x, float* y, int Bn, int H, int W, int C, int KH, int KW, int stride, int pad, int OH,
                           int OW, cudaStream_t s);
cudaError_t launch_avgpool(const float* x, float* y, int Bn, int HW, int C, cudaStream_t s);
// y = LayerNorm(x (+res)) or, with ids != nullptr, LayerNorm(word[id] + pos[s] + type[0]) (BERT embeddings)
cudaError_t launch_layernorm(const float* x, const float* res, const int* ids, const float* word, const float* pos,
                             const float* type, const float* gamma, const float* beta, float* y, int tokens, int S, int H,
                             int vocab, float eps, cudaStream_t s);
cudaError_t launch_attention(const float* qkv, const int* ids, float* ctx, int Bn, int S, int H, int heads, cudaStream_t s);
size_t attention_smem_bytes(int S, int H, int heads);

// tcgen05 3xTF32 version of launch_gemm (gemm_tc.cu) for M >= 64, N % 32 == 0, K >= 32, lda % 4 == 0
bool gemm_tc_supported(const float* A, const float* B, const float* bias, const float* R, const float* C, int M, int N, int K,
                       int lda);
cudaError_t launch_gemm_tc(const float* A, const float* B, const float* bias, const float* R, float* C, int M, int N, int K,
                           int lda, int act, cudaStream_t s);

// X6 (+ X7): batch gather / scatter as ONE kernel over a table of (src, dst, bytes) segments. A source / destination may be
// pinned host memory (zero-copy over PCIe: the client thread wrote it, no batcher-side memcpy, no staging copy), local HBM,
// or another GPU's forward window mapped through CUDA IPC (the forward hop a6: NVLink loads / stores inside this kernel).
struct CopySeg {
  const void* src;
  void* dst;
  uint64_t bytes;
};
cudaError_t launch_copy_segments(const CopySeg* segs, int n, cudaStream_t s);

// X4: implicit-GEMM convolution on tcgen05 (gemm_tc.cu): y[B,OH,OW,N] = act(conv(x[B,H,W,C], w[KH,KW,C,N]) + bias (+ R)); the A
// tiles are gathered from the NHWC activations by TMA im2col tensor maps -- no patch matrix in HBM. C % 32 == 0, N % 32 == 0.
bool conv_tc_supported(const float* x, const float* w, const float* bias, const float* R, const float* y, int Bn, int H, int W,
                       int C, int KH, int KW, int stride, int pad, int OH, int OW, int N);
cudaError_t launch_conv_tc(const float* x, const float* w, const float* bias, const float* R, float* y, int Bn, int H, int W, int C,
                           int KH, int KW, int stride, int pad, int OH, int OW, int N, int act, cudaStream_t s);

int gemm_trace_read(long long* out16);  // debugging aid (TFSC_GT_TRACE=1): clock64 timeline of CTA 0 of the last persistent GEMM

int64_t kernel_launch_count();

}  // namespace tfsc
