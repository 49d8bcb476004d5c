// tests/hostcheck.cu -- runs the product's __host__ __device__ hash / partition
// functions (dpark_b200/csrc/dpk_common.cuh) on the CPU so that the arithmetic
// can be checked against the oracle without a GPU.  Test-only; not shipped.
#include "dpk_common.cuh"
extern "C" {
void hc_hash_i64(const int64_t *k, int64_t n, int64_t *o) { for (int64_t i = 0; i < n; i++) o[i] = dpk::hash_i64(k[i]); }
void hc_hash_u64(const uint64_t *k, int64_t n, int64_t *o) { for (int64_t i = 0; i < n; i++) o[i] = dpk::hash_u64(k[i]); }
void hc_hash_f64(const double *k, int64_t n, int64_t *o) { for (int64_t i = 0; i < n; i++) o[i] = dpk::hash_f64(k[i]); }
void hc_hash_bytes(const uint8_t *d, const int64_t *off, int64_t n, int mode, int64_t *o) {
    for (int64_t i = 0; i < n; i++)
        o[i] = mode == 0 ? dpk::hash_bytes_signed(d + off[i], off[i + 1] - off[i])
                         : dpk::hash_utf8_codepoints(d + off[i], off[i + 1] - off[i]);
}
void hc_hash_tuple(const int64_t *item_hash, int64_t n, int32_t arity, int64_t *o) {
    for (int64_t i = 0; i < n; i++) o[i] = dpk::hash_tuple_items(item_hash, n, i, arity);
}
int hc_partition(const int64_t *h, int64_t n, int32_t P, const int64_t *thr, int32_t nthr, int32_t *o) {
    dpk::PartFn f;
    int rc = dpk::make_partfn(P, thr, nthr, 0, &f);
    if (rc) return rc;
    for (int64_t i = 0; i < n; i++) o[i] = f(h[i]);
    return 0;
}
int hc_bucket(const int64_t *h, int64_t n, int32_t P, int32_t sub_bits, int32_t *o) {
    dpk::PartFn f;
    int rc = dpk::make_partfn(P, nullptr, 0, sub_bits, &f);
    if (rc) return rc;
    for (int64_t i = 0; i < n; i++) o[i] = f.bucket(h[i]);
    return 0;
}
}
