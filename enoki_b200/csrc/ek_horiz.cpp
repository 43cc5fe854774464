/*
 * ek_horiz.cpp -- eager horizontal operations on raw device memory.
 *
 * ABI-compatible stand-ins for cuda_hsum/hprod/hmax/hmin/count/all/any
 * (src/cuda/horiz.cu:162-354).  They are thin wrappers over the evaluator's fused
 * reduction epilogue (no CUB, no extra temporary): the raw pointer is wrapped in a
 * borrowed variable, a lazy reduction node is recorded and evaluated, and the
 * 8-byte result buffer is handed to the caller (who owns it, like the reference).
 */
#include "ek_internal.h"
#include <cstring>

static void *reduce_raw(ek_type type, ek_type rtype, ek_op op, size_t n, const void *data) {
    if (n == 0) { ek_set_error("horizontal reduction of an empty array"); return nullptr; }
    uint32_t src = ek_var_register(type, n, const_cast<void *>(data), 0);
    if (!src) return nullptr;
    uint32_t r = ek_trace_append(rtype, op, src, 0, 0, 0);
    ek_dec_ref_ext(src);
    if (!r) return nullptr;
    if (ek_eval_var(r) != 0) { ek_dec_ref_ext(r); return nullptr; }
    EkVariable &v = ek_ctx().vars[r];
    void *out = v.data;
    v.data = nullptr; v.free_data = false;      /* ownership moves to the caller */
    ek_dec_ref_ext(r);
    return out;
}

extern "C" {

void *ek_hsum(ek_type type, size_t n, const void *data)  { return reduce_raw(type, type, EK_OP_HSUM, n, data); }
void *ek_hprod(ek_type type, size_t n, const void *data) { return reduce_raw(type, type, EK_OP_HPROD, n, data); }
void *ek_hmax(ek_type type, size_t n, const void *data)  { return reduce_raw(type, type, EK_OP_HMAX, n, data); }
void *ek_hmin(ek_type type, size_t n, const void *data)  { return reduce_raw(type, type, EK_OP_HMIN, n, data); }

static uint32_t fetch_u32(void *p) {
    uint32_t v = 0;
    ek_memcpy_from_device(&v, p, 4);        /* blocking D2H like horiz.cu:303,328,350 */
    ek_free(p);
    return v;
}
size_t ek_count(size_t n, const uint8_t *mask) {
    void *p = reduce_raw(EK_BOOL, EK_UINT32, EK_OP_COUNT, n, mask);
    return p ? fetch_u32(p) : 0;
}
int ek_all(size_t n, const uint8_t *mask) {
    void *p = reduce_raw(EK_BOOL, EK_BOOL, EK_OP_ALL, n, mask);
    return p ? (fetch_u32(p) & 0xff) != 0 : 0;
}
int ek_any(size_t n, const uint8_t *mask) {
    void *p = reduce_raw(EK_BOOL, EK_BOOL, EK_OP_ANY, n, mask);
    return p ? (fetch_u32(p) & 0xff) != 0 : 0;
}

} /* extern "C" */
