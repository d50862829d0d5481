/* dflash_b200 — C ABI of the B200-native DFlash draft-and-verify hot path.
 *
 * Plain pointers and sizes only (no torch types). Every pointer is a DEVICE pointer unless the
 * name ends in _host. `stream` is a cudaStream_t passed as void*. No entry point allocates device
 * memory or synchronises; all return DFLASH_OK (0) or a negative error (dflash_last_error() holds
 * the message for the calling thread). The library contains sm_100a code only.
 *
 * The reference (AtharvRN/dflash) is pure Python; it has no FFI. Each entry point names the
 * reference call site(s) (file:line under the reference tree) it replaces; INTEGRATION.md shows
 * the ctypes binding a reference maintainer would add.
 */
#ifndef DFLASH_B200_H_
#define DFLASH_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define DFLASH_ABI_VERSION 1

#define DFLASH_OK 0
#define DFLASH_ERR_ARG (-1)   /* bad argument / unsupported shape */
#define DFLASH_ERR_CUDA (-2)  /* CUDA runtime / driver call failed */
#define DFLASH_ERR_ARCH (-3)  /* not an sm_100a device */

int dflash_abi_version(void);
const char* dflash_last_error(void);

/* Returns the SM count of the current device (>0) or DFLASH_ERR_ARCH if it is not sm_100. */
int dflash_device_check(void);

/* ---------------------------------------------------------------------------------------------
 * Raw operators (unit-test granularity).
 * ------------------------------------------------------------------------------------------- */

/* out[m, n] = sum_k X[x_row0+m, k] * W[w_row0+n, k]   (fp32 out; bf16 in; nn.Linear layout)
 * Replaces torch.nn.Linear on the draft path (model/dflash.py:70-76,101,177; Qwen3MLP via :143).
 * mb = padded row count fed to the tensor core (16/32/64/128/256), m_valid <= mb rows are written.
 * ws: fp32 scratch [dflash_gemm_max_slots(N,K,grid)][ws_rows][ws_ld] for the split-K partials. */
int dflash_gemm_max_slots(int N, int K, int grid);
int dflash_gemm_skinny(const void* W, int w_rows_total, int w_row0, int N, int K, const void* X,
                       int x_rows_total, int x_row0, int mb, int m_valid, float* ws, int ws_rows,
                       long long ws_ld, float* out, long long out_ld, int grid, int use_pdl,
                       void* stream);

/* tokens_out[m] = argmax_n bf16(sum_k X[x_row0+m,k] * W[n,k]), ties -> lowest n; optionally also
 * writes the bf16 logits. Replaces target.lm_head(...) + sample(draft_logits)
 * (model/dflash.py:238-247, model/utils.py:27-29) without materialising the logits.
 * cand_val/cand_idx: scratch [grid][mb]. */
int dflash_gemm_argmax(const void* W, int w_rows_total, int N, int K, const void* X, int x_rows_total,
                       int x_row0, int mb, int m_valid, float* cand_val, int* cand_idx, void* logits,
                       long long logits_ld, long long* tokens_out, int grid, int use_pdl, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DFLASH_B200_H_ */
