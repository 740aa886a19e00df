// capi.cu -- extern "C" entry points of liblac_b200.so (include/lac_b200.h).
// Argument checking, launch, error text.  No CPU compute path exists here: every entry
// point either launches CUDA work or fails with LAC_E_CUDA.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <cuda_runtime.h>

#include "../../include/lac_b200.h"

namespace lac {
cudaError_t launch_lookup(const float*, int64_t, int, int64_t, const int32_t*, uint32_t*, uint32_t*, cudaStream_t);
cudaError_t launch_build(const float*, int64_t, int, int64_t, uint32_t*, cudaStream_t);
cudaError_t launch_uniform_encode(const int32_t*, int64_t, int64_t, int64_t, const int32_t*, int, lac_enc_state*, uint8_t*,
                                  int64_t, int, int, cudaStream_t);
cudaError_t launch_uniform_decode(int64_t, int64_t, const int32_t*, int, lac_dec_state*, const uint8_t*, const int64_t*,
                                  int32_t*, int64_t, int, cudaStream_t);
cudaError_t launch_decode(const float*, int64_t, int64_t, int64_t, int64_t, int, const int32_t*, lac_dec_state*,
                          const uint8_t*, const int64_t*, int32_t*, int64_t, int, cudaStream_t);
cudaError_t launch_dec_init(lac_dec_state*, int64_t, int, const uint8_t*, const int64_t*, cudaStream_t);
int max_vocab_single_cta();
int max_vocab();
cudaError_t launch_enc_init(lac_enc_state*, int64_t, int, cudaStream_t);
cudaError_t launch_encode_pairs(const uint32_t*, int64_t, int64_t, int64_t, int64_t, const int32_t*, lac_enc_state*,
                                uint8_t*, int64_t, int, int, cudaStream_t);
cudaError_t launch_ac_tables_encode(const int64_t*, int, int64_t, int64_t, const int64_t*, int64_t, int64_t,
                                    const int32_t*, int64_t, int64_t, const int32_t*, lac_enc_state*, uint8_t*,
                                    int64_t, int, int, int, cudaStream_t);
cudaError_t launch_ac_tables_decode(const int64_t*, int, int64_t, int64_t, const int64_t*, int64_t, int64_t, int64_t,
                                    int64_t, const int32_t*, lac_dec_state*, const uint8_t*, const int64_t*,
                                    int32_t*, int64_t, int, int, cudaStream_t);
cudaError_t launch_acs_tables_encode(const uint64_t*, int, int64_t, int64_t, const int32_t*, int64_t, int64_t,
                                     const int32_t*, lac_enc_state*, uint8_t*, int64_t, int, int, cudaStream_t);
cudaError_t launch_acs_tables_decode(const uint64_t*, int, int64_t, int64_t, int64_t, int64_t, const int32_t*,
                                     lac_dec_state*, const uint8_t*, const int64_t*, int32_t*, int64_t, int,
                                     cudaStream_t);
}  // namespace lac

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    return fail(LAC_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
}
#define CK(call, what)                                  \
    do {                                                \
        cudaError_t e__ = (call);                       \
        if (e__ != cudaSuccess) return cuda_fail(e__, what); \
    } while (0)

bool prec_ok(int prec, int lo) { return prec >= lo && prec <= 60; }

int check_vocab(int32_t vocab, const void* logits = nullptr, int64_t s0 = 0, int64_t s1 = 0) {
    if (vocab < 1 || vocab > lac::max_vocab())
        return fail(LAC_E_ARG, "vocab %d out of range [1, %d]", vocab, lac::max_vocab());
    if (vocab > lac::max_vocab_single_cta()) {
        // rows wider than one CTA's registers are split over a thread-block cluster, TMA-staged only
        if (vocab % 4 != 0 || (((uintptr_t)logits) & 15) != 0 || s0 % 4 != 0 || s1 % 4 != 0)
            return fail(LAC_E_ARG, "vocab %d > %d needs 16-byte aligned rows (vocab and strides multiples of 4)", vocab,
                        lac::max_vocab_single_cta());
    }
    return LAC_OK;
}

}  // namespace

extern "C" {

int lac_abi_version(void) { return LAC_ABI_VERSION; }
const char* lac_last_error(void) { return g_err.c_str(); }

int lac_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* hbm_bytes) {
    int dev = 0;
    CK(cudaGetDevice(&dev), "cudaGetDevice");
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, dev), "cudaGetDeviceProperties");
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (hbm_bytes) *hbm_bytes = (int64_t)p.totalGlobalMem;
    return LAC_OK;
}

int lac_cdf_build_f32(const float* d_logits, int64_t rows, int32_t vocab, int64_t row_stride, uint32_t* d_cum,
                      void* stream) {
    if (!d_logits || !d_cum || rows < 0 || row_stride < vocab) return fail(LAC_E_ARG, "lac_cdf_build_f32: bad argument");
    if (int rc = check_vocab(vocab, d_logits, row_stride)) return rc;
    CK(lac::launch_build(d_logits, rows, vocab, row_stride, d_cum, (cudaStream_t)stream), "lac_cdf_build_f32");
    return LAC_OK;
}

int lac_cdf_lookup_f32(const float* d_logits, int64_t rows, int32_t vocab, int64_t row_stride, const int32_t* d_syms,
                       uint32_t* d_pairs, uint32_t* d_status, void* stream) {
    if (!d_logits || !d_syms || !d_pairs || rows < 0 || row_stride < vocab)
        return fail(LAC_E_ARG, "lac_cdf_lookup_f32: bad argument");
    if (int rc = check_vocab(vocab, d_logits, row_stride)) return rc;
    if ((uintptr_t)d_pairs & 7) return fail(LAC_E_ARG, "lac_cdf_lookup_f32: d_pairs must be 8-byte aligned");
    CK(lac::launch_lookup(d_logits, rows, vocab, row_stride, d_syms, d_pairs, d_status, (cudaStream_t)stream),
       "lac_cdf_lookup_f32");
    return LAC_OK;
}

int lac_enc_init(lac_enc_state* d_state, int64_t n_streams, int prec, void* stream) {
    if (!d_state || n_streams < 0 || !prec_ok(prec, 2)) return fail(LAC_E_ARG, "lac_enc_init: bad argument");
    CK(lac::launch_enc_init(d_state, n_streams, prec, (cudaStream_t)stream), "lac_enc_init");
    return LAC_OK;
}

int lac_dec_init(lac_dec_state* d_state, int64_t n_streams, int prec, const uint8_t* d_bytes,
                 const int64_t* d_offsets, void* stream) {
    if (!d_state || !d_bytes || !d_offsets || n_streams < 0 || !prec_ok(prec, 2))
        return fail(LAC_E_ARG, "lac_dec_init: bad argument");
    CK(lac::launch_dec_init(d_state, n_streams, prec, d_bytes, d_offsets, (cudaStream_t)stream), "lac_dec_init");
    return LAC_OK;
}

int lac_ac_encode_pairs(const uint32_t* d_pairs, int64_t n_streams, int64_t T, int64_t stream_stride,
                        int64_t tok_stride, const int32_t* d_ntok, lac_enc_state* d_state, uint8_t* d_out,
                        int64_t out_stride, int finish, int prec, void* stream) {
    if ((!d_pairs && T > 0) || !d_state || !d_out || n_streams < 0 || T < 0 || out_stride < 1)
        return fail(LAC_E_ARG, "lac_ac_encode_pairs: bad argument");
    if (!prec_ok(prec, 34)) return fail(LAC_E_ARG, "lac_ac_encode_pairs: prec %d outside [34, 60]", prec);
    if ((uintptr_t)d_pairs & 7) return fail(LAC_E_ARG, "lac_ac_encode_pairs: d_pairs must be 8-byte aligned");
    CK(lac::launch_encode_pairs(d_pairs, n_streams, T, stream_stride, tok_stride, d_ntok, d_state, d_out, out_stride,
                                finish, prec, (cudaStream_t)stream),
       "lac_ac_encode_pairs");
    return LAC_OK;
}

int lac_ac_decode_logits_f32(const float* d_logits, int64_t n_streams, int64_t T, int64_t stream_stride,
                             int64_t tok_stride, int32_t vocab, const int32_t* d_ntok, lac_dec_state* d_state,
                             const uint8_t* d_bytes, const int64_t* d_offsets, int32_t* d_syms, int64_t sym_stride,
                             int prec, void* stream) {
    if ((!d_logits && T > 0) || !d_state || !d_bytes || !d_offsets || (!d_syms && T > 0) || n_streams < 0 || T < 0)
        return fail(LAC_E_ARG, "lac_ac_decode_logits_f32: bad argument");
    if (!prec_ok(prec, 34)) return fail(LAC_E_ARG, "lac_ac_decode_logits_f32: prec %d outside [34, 60]", prec);
    if (int rc = check_vocab(vocab, d_logits, stream_stride, tok_stride)) return rc;
    CK(lac::launch_decode(d_logits, n_streams, T, stream_stride, tok_stride, vocab, d_ntok, d_state, d_bytes,
                          d_offsets, d_syms, sym_stride, prec, (cudaStream_t)stream),
       "lac_ac_decode_logits_f32");
    return LAC_OK;
}

int lac_ac_encode_uniform(const int32_t* d_syms, int64_t n_streams, int64_t T, int64_t sym_stride,
                          const int32_t* d_ntok, int32_t n_symbols, lac_enc_state* d_state, uint8_t* d_out,
                          int64_t out_stride, int finish, int prec, void* stream) {
    if ((!d_syms && T > 0) || !d_state || !d_out || n_streams < 0 || T < 0 || out_stride < 1 || n_symbols < 1)
        return fail(LAC_E_ARG, "lac_ac_encode_uniform: bad argument");
    if (!prec_ok(prec, 2)) return fail(LAC_E_ARG, "lac_ac_encode_uniform: prec %d outside [2, 60]", prec);
    CK(lac::launch_uniform_encode(d_syms, n_streams, T, sym_stride, d_ntok, n_symbols, d_state, d_out, out_stride,
                                  finish, prec, (cudaStream_t)stream),
       "lac_ac_encode_uniform");
    return LAC_OK;
}

int lac_ac_decode_uniform(int64_t n_streams, int64_t T, const int32_t* d_ntok, int32_t n_symbols,
                          lac_dec_state* d_state, const uint8_t* d_bytes, const int64_t* d_offsets, int32_t* d_syms,
                          int64_t sym_stride, int prec, void* stream) {
    if (!d_state || !d_bytes || !d_offsets || (!d_syms && T > 0) || n_streams < 0 || T < 0 || n_symbols < 1)
        return fail(LAC_E_ARG, "lac_ac_decode_uniform: bad argument");
    if (!prec_ok(prec, 2)) return fail(LAC_E_ARG, "lac_ac_decode_uniform: prec %d outside [2, 60]", prec);
    CK(lac::launch_uniform_decode(n_streams, T, d_ntok, n_symbols, d_state, d_bytes, d_offsets, d_syms, sym_stride,
                                  prec, (cudaStream_t)stream),
       "lac_ac_decode_uniform");
    return LAC_OK;
}

int lac_ac_encode_tables(const int64_t* d_dist, int32_t vocab, int64_t stream_stride, int64_t tok_stride,
                         const int64_t* d_minp, int64_t minp_stream_stride, int64_t minp_tok_stride,
                         const int32_t* d_syms, int64_t n_streams, int64_t T, const int32_t* d_ntok,
                         lac_enc_state* d_state, uint8_t* d_out, int64_t out_stride, int finish, int prec, int flags,
                         void* stream) {
    if (!d_dist || !d_minp || (!d_syms && T > 0) || !d_state || !d_out || vocab < 1 || n_streams < 0 || T < 0 ||
        out_stride < 1)
        return fail(LAC_E_ARG, "lac_ac_encode_tables: bad argument");
    if (!prec_ok(prec, 2)) return fail(LAC_E_ARG, "lac_ac_encode_tables: prec %d outside [2, 60]", prec);
    CK(lac::launch_ac_tables_encode(d_dist, vocab, stream_stride, tok_stride, d_minp, minp_stream_stride,
                                    minp_tok_stride, d_syms, n_streams, T, d_ntok, d_state, d_out, out_stride, finish,
                                    prec, flags, (cudaStream_t)stream),
       "lac_ac_encode_tables");
    return LAC_OK;
}

int lac_ac_decode_tables(const int64_t* d_dist, int32_t vocab, int64_t stream_stride, int64_t tok_stride,
                         const int64_t* d_minp, int64_t minp_stream_stride, int64_t minp_tok_stride,
                         int64_t n_streams, int64_t T, const int32_t* d_ntok, lac_dec_state* d_state,
                         const uint8_t* d_bytes, const int64_t* d_offsets, int32_t* d_syms, int64_t sym_stride,
                         int prec, int flags, void* stream) {
    if (!d_dist || !d_minp || !d_state || !d_bytes || !d_offsets || !d_syms || vocab < 1 || n_streams < 0 || T < 0)
        return fail(LAC_E_ARG, "lac_ac_decode_tables: bad argument");
    if (!prec_ok(prec, 2)) return fail(LAC_E_ARG, "lac_ac_decode_tables: prec %d outside [2, 60]", prec);
    CK(lac::launch_ac_tables_decode(d_dist, vocab, stream_stride, tok_stride, d_minp, minp_stream_stride,
                                    minp_tok_stride, n_streams, T, d_ntok, d_state, d_bytes, d_offsets, d_syms,
                                    sym_stride, prec, flags, (cudaStream_t)stream),
       "lac_ac_decode_tables");
    return LAC_OK;
}

int lac_acs_encode_tables(const uint64_t* d_cdf, int32_t vocab, int64_t stream_stride, int64_t tok_stride,
                          const int32_t* d_syms, int64_t n_streams, int64_t T, const int32_t* d_ntok,
                          lac_enc_state* d_state, uint8_t* d_out, int64_t out_stride, int finish, int prec,
                          void* stream) {
    if (!d_cdf || (!d_syms && T > 0) || !d_state || !d_out || vocab < 1 || n_streams < 0 || T < 0 || out_stride < 1)
        return fail(LAC_E_ARG, "lac_acs_encode_tables: bad argument");
    if (!prec_ok(prec, 2)) return fail(LAC_E_ARG, "lac_acs_encode_tables: prec %d outside [2, 60]", prec);
    CK(lac::launch_acs_tables_encode(d_cdf, vocab, stream_stride, tok_stride, d_syms, n_streams, T, d_ntok, d_state,
                                     d_out, out_stride, finish, prec, (cudaStream_t)stream),
       "lac_acs_encode_tables");
    return LAC_OK;
}

int lac_acs_decode_tables(const uint64_t* d_cdf, int32_t vocab, int64_t stream_stride, int64_t tok_stride,
                          int64_t n_streams, int64_t T, const int32_t* d_ntok, lac_dec_state* d_state,
                          const uint8_t* d_bytes, const int64_t* d_offsets, int32_t* d_syms, int64_t sym_stride,
                          int prec, void* stream) {
    if (!d_cdf || !d_state || !d_bytes || !d_offsets || !d_syms || vocab < 1 || n_streams < 0 || T < 0)
        return fail(LAC_E_ARG, "lac_acs_decode_tables: bad argument");
    if (!prec_ok(prec, 2)) return fail(LAC_E_ARG, "lac_acs_decode_tables: prec %d outside [2, 60]", prec);
    CK(lac::launch_acs_tables_decode(d_cdf, vocab, stream_stride, tok_stride, n_streams, T, d_ntok, d_state, d_bytes,
                                     d_offsets, d_syms, sym_stride, prec, (cudaStream_t)stream),
       "lac_acs_decode_tables");
    return LAC_OK;
}

// ------------------------------------------------------------------ host-buffer calls
namespace {

struct DevBuf {
    void* p = nullptr;
    size_t n = 0;
    int reserve(size_t bytes) {
        if (bytes <= n) return LAC_OK;
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
        n = bytes;
        return LAC_OK;
    }
};

struct HostCtx {
    std::mutex mu;
    cudaStream_t st = nullptr;
    DevBuf logits, syms, pairs, state, out, offs, bytes;
    int ensure_stream() {
        if (!st) CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking), "cudaStreamCreate");
        return LAC_OK;
    }
};
HostCtx g_host;

constexpr size_t kSliceBytes = size_t(1) << 30;  // logits staged to the device in <= 1 GiB slices

}  // namespace

int lac_encode_logits_host(const float* h_logits, const int32_t* h_syms, int64_t n_streams, int64_t T,
                           int32_t vocab, uint8_t* h_out, int64_t out_stride, uint64_t* h_nbits, int prec) {
    if (!h_logits || !h_syms || !h_out || !h_nbits || n_streams < 0 || T < 0 || out_stride < 1)
        return fail(LAC_E_ARG, "lac_encode_logits_host: bad argument");
    if (!prec_ok(prec, 34)) return fail(LAC_E_ARG, "lac_encode_logits_host: prec %d outside [34, 60]", prec);
    if (int rc = check_vocab(vocab)) return rc;
    std::lock_guard<std::mutex> lk(g_host.mu);
    if (int rc = g_host.ensure_stream()) return rc;
    cudaStream_t st = g_host.st;
    const int64_t rows = n_streams * T;
    if (rows == 0) return LAC_OK;
    const size_t row_bytes = (size_t)vocab * 4;
    int64_t slice_rows = (int64_t)(kSliceBytes / row_bytes);
    if (slice_rows < 1) slice_rows = 1;
    if (slice_rows > rows) slice_rows = rows;
    if (int rc = g_host.logits.reserve((size_t)slice_rows * row_bytes)) return rc;
    if (int rc = g_host.syms.reserve((size_t)rows * 4)) return rc;
    if (int rc = g_host.pairs.reserve((size_t)rows * 8)) return rc;
    if (int rc = g_host.state.reserve((size_t)n_streams * sizeof(lac_enc_state))) return rc;
    if (int rc = g_host.out.reserve((size_t)n_streams * (size_t)out_stride)) return rc;
    CK(cudaMemcpyAsync(g_host.syms.p, h_syms, (size_t)rows * 4, cudaMemcpyHostToDevice, st), "H2D syms");
    for (int64_t r0 = 0; r0 < rows; r0 += slice_rows) {
        int64_t nr = rows - r0 < slice_rows ? rows - r0 : slice_rows;
        CK(cudaMemcpyAsync(g_host.logits.p, h_logits + r0 * vocab, (size_t)nr * row_bytes, cudaMemcpyHostToDevice, st),
           "H2D logits");
        CK(lac::launch_lookup((const float*)g_host.logits.p, nr, vocab, vocab, (const int32_t*)g_host.syms.p + r0,
                              (uint32_t*)g_host.pairs.p + 2 * r0, nullptr, st),
           "lookup");
    }
    lac_enc_state* dstate = (lac_enc_state*)g_host.state.p;
    CK(lac::launch_enc_init(dstate, n_streams, prec, st), "enc_init");
    CK(lac::launch_encode_pairs((const uint32_t*)g_host.pairs.p, n_streams, T, T, 1, nullptr, dstate,
                                (uint8_t*)g_host.out.p, out_stride, 1, prec, st),
       "encode_pairs");
    CK(cudaMemcpyAsync(h_out, g_host.out.p, (size_t)n_streams * (size_t)out_stride, cudaMemcpyDeviceToHost, st),
       "D2H bytes");
    std::string tmp((size_t)n_streams * sizeof(lac_enc_state), '\0');
    CK(cudaMemcpyAsync(&tmp[0], dstate, tmp.size(), cudaMemcpyDeviceToHost, st), "D2H state");
    CK(cudaStreamSynchronize(st), "sync");
    const lac_enc_state* hs = (const lac_enc_state*)tmp.data();
    int rc = LAC_OK;
    for (int64_t s = 0; s < n_streams; s++) {
        h_nbits[s] = hs[s].nbits;
        if (hs[s].status & LAC_ST_CAP) rc = fail(LAC_E_CAP, "stream %lld: output capacity exceeded", (long long)s);
        else if (hs[s].status) rc = fail(LAC_E_SYMBOL, "stream %lld: status %u", (long long)s, hs[s].status);
    }
    return rc;
}

int lac_decode_logits_host(const float* h_logits, int64_t n_streams, int64_t T, int32_t vocab,
                           const uint8_t* h_bytes, const int64_t* h_offsets, int32_t* h_syms, int prec) {
    if (!h_logits || !h_bytes || !h_offsets || !h_syms || n_streams < 0 || T < 0)
        return fail(LAC_E_ARG, "lac_decode_logits_host: bad argument");
    if (!prec_ok(prec, 34)) return fail(LAC_E_ARG, "lac_decode_logits_host: prec %d outside [34, 60]", prec);
    if (int rc = check_vocab(vocab)) return rc;
    std::lock_guard<std::mutex> lk(g_host.mu);
    if (int rc = g_host.ensure_stream()) return rc;
    cudaStream_t st = g_host.st;
    if (n_streams == 0 || T == 0) return LAC_OK;
    const size_t stream_bytes = (size_t)T * (size_t)vocab * 4;
    int64_t slice = (int64_t)(kSliceBytes / stream_bytes);
    if (slice < 1) slice = 1;
    if (slice > n_streams) slice = n_streams;
    const size_t total_bytes = (size_t)h_offsets[n_streams];
    if (int rc = g_host.logits.reserve((size_t)slice * stream_bytes)) return rc;
    if (int rc = g_host.bytes.reserve(total_bytes + 16)) return rc;
    if (int rc = g_host.offs.reserve((size_t)(n_streams + 1) * 8)) return rc;
    if (int rc = g_host.state.reserve((size_t)n_streams * sizeof(lac_dec_state))) return rc;
    if (int rc = g_host.syms.reserve((size_t)n_streams * (size_t)T * 4)) return rc;
    CK(cudaMemcpyAsync(g_host.bytes.p, h_bytes, total_bytes, cudaMemcpyHostToDevice, st), "H2D bytes");
    CK(cudaMemcpyAsync(g_host.offs.p, h_offsets, (size_t)(n_streams + 1) * 8, cudaMemcpyHostToDevice, st), "H2D offsets");
    lac_dec_state* dstate = (lac_dec_state*)g_host.state.p;
    CK(lac::launch_dec_init(dstate, n_streams, prec, (const uint8_t*)g_host.bytes.p, (const int64_t*)g_host.offs.p, st),
       "dec_init");
    for (int64_t s0 = 0; s0 < n_streams; s0 += slice) {
        int64_t ns = n_streams - s0 < slice ? n_streams - s0 : slice;
        CK(cudaMemcpyAsync(g_host.logits.p, h_logits + s0 * T * vocab, (size_t)ns * stream_bytes,
                           cudaMemcpyHostToDevice, st),
           "H2D logits");
        CK(lac::launch_decode((const float*)g_host.logits.p, ns, T, T * (int64_t)vocab, vocab, vocab, nullptr,
                              dstate + s0, (const uint8_t*)g_host.bytes.p, (const int64_t*)g_host.offs.p + s0,
                              (int32_t*)g_host.syms.p + s0 * T, T, prec, st),
           "decode");
    }
    CK(cudaMemcpyAsync(h_syms, g_host.syms.p, (size_t)n_streams * (size_t)T * 4, cudaMemcpyDeviceToHost, st),
       "D2H syms");
    CK(cudaStreamSynchronize(st), "sync");
    return LAC_OK;
}

}  // extern "C"
