/*
 * Debug / probe entry points of the ucd_b200 kernels.  NOT part of the product boundary: they are exported only by
 * ucd_b200/libucd_b200_debug.so (the same sources compiled with -DUCD_DEBUG_KNOBS, `python -m ucd_b200.build --debug`),
 * which is also the only build that keeps global debug state (the trace pointer) and reads environment variables
 * (UCD_SPLITS1, UCD_SPLITS2, UCD_UP_GY, UCD_UPB_UN, UCD_UPB_BPS tuning knobs).  Used by scripts/ and by the tcgen05 building-block self-test.
 */
#ifndef UCD_B200_DEBUG_H
#define UCD_B200_DEBUG_H
#include "ucd_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* per-role cycle counters of the sweep CTAs, see contrast.cu */
int ucd_con_debug_trace(void* device_buffer_or_null);
int ucd_con_debug_splits(int64_t max_row_tiles, int64_t max_col_tiles);

/* Self-test of the tcgen05/TMEM/bulk-copy building blocks: C[M=128,N] = A[128,K] * B[N,K]^T in bf16
 * with the production tile layout, and D[128,256] = E[128,128] * Bt (MN-major B). Returns max abs
 * error through *max_err_host (synchronises). */
int ucd_selftest_umma(int variant, float* max_err_host);
/* tcgen05.mma issue-rate probe (cycles per instruction for a chain of `iters` MMAs; modes in selftest.cu) */
int ucd_selftest_mma_rate(int mode, int iters, float* cycles_per_instr_host);
/* CUDA-core pipe probe behind the sweep epilogue's design (ex2 / bf16 pack rates; modes in selftest.cu) */
int ucd_selftest_pipe_rate(int mode, int warps, int iters, float* cycles_per_iter_host);

/* tensor-pipe time per column tile of the sweep-1 MMA mix under a given tensor-memory placement (selftest.cu) */
int ucd_selftest_mma_mix(int s_ts, int s_a, int s_acc0, int s_acc1, int v_n256, int v_a, int v_acc, int tiles,
                         float* cycles_per_tile_host);

/* HBM read-only probe (selftest.cu): microseconds per pass over `bytes` of `buf`; mode 0 register loads, 1 cp.async
 * ring, 2 strided NCHW walk like the loss kernels; `un` loads in flight per thread, `blocks_per_sm` blocks of 256 */
int ucd_selftest_read_probe(const void* buf, long long bytes, int mode, int un, int blocks_per_sm, int reps,
                            float* us_host);

#ifdef __cplusplus
}
#endif
#endif /* UCD_B200_DEBUG_H */
