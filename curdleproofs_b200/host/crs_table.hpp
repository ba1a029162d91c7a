// Process-wide cache of CRS digit tables (cdp_fixed_table, include/cdp_msm.h).  A prover and a verifier over the same CRS
// (/root/reference/src/crs.rs:19-34) on the same device -- or several provers with different batch sizes -- share ONE table
// (12 GiB at ell = 252 with 16-bit windows) instead of building one each; the table is immutable after creation and may be read from any
// context of its device concurrently.  Reference-counted; the last release frees it.
#pragma once
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/cdp_msm.h"

namespace cdp_host {

struct SharedCrsTable {
    cdp_fixed_table *table = nullptr;
    cdp_ctx *ctx = nullptr;          // private context the table was built on (outlives every user's context)
    std::vector<uint8_t> crs_ext;    // vec_G | vec_H | H | G_t | G_u | sum(G) | sum(Hvec): ell + 9 affine points, the table's base order
    size_t ell = 0;
    int device = 0, bits = 0, refs = 0;
};

inline std::mutex &crs_table_mutex() { static std::mutex m; return m; }
inline std::vector<SharedCrsTable *> &crs_table_list() { static std::vector<SharedCrsTable *> v; return v; }

// crs_points: ell + 7 affine points in `CurdleproofsCrs::from_points` order (src/crs.rs:37-58).  CDP_FIXED_BITS overrides the window width.
inline int crs_table_acquire(cdp_ctx *user_ctx, size_t ell, const uint8_t *crs_points, SharedCrsTable **out) {
    constexpr size_t NBL = 4;
    *out = nullptr;
    int bits = 0;
    if (const char *e = getenv("CDP_FIXED_BITS")) bits = atoi(e);
    const int device = cdp_ctx_device(user_ctx);
    std::lock_guard<std::mutex> lock(crs_table_mutex());
    for (SharedCrsTable *t : crs_table_list())
        if (t->device == device && t->bits == bits && t->ell == ell && memcmp(t->crs_ext.data(), crs_points, (ell + 7) * 96) == 0) {
            t->refs++;
            *out = t;
            return CDP_OK;
        }
    SharedCrsTable *t = new SharedCrsTable();
    t->ell = ell; t->device = device; t->bits = bits; t->refs = 1;
    int rc = cdp_ctx_create(&t->ctx, device, nullptr);
    if (rc != CDP_OK) { delete t; return rc; }
    t->crs_ext.resize((ell + 9) * 96);
    memcpy(t->crs_ext.data(), crs_points, (ell + 7) * 96);
    // crs.G_sum / crs.H_sum (src/crs.rs:46-47) as two more table bases
    const size_t no = ell > NBL ? ell : NBL;
    std::vector<uint8_t> ones(32 * no, 0), sums(2 * 144);
    for (size_t i = 0; i < no; i++) ones[32 * i] = 1;
    rc = cdp_msm(t->ctx, crs_points, ones.data(), ell, sums.data());
    if (rc == CDP_OK) rc = cdp_msm(t->ctx, crs_points + ell * 96, ones.data(), NBL, sums.data() + 144);
    if (rc == CDP_OK) rc = cdp_normalize_batch(t->ctx, sums.data(), 2, t->crs_ext.data() + (ell + 7) * 96);
    if (rc == CDP_OK) rc = cdp_fixed_table_create(t->ctx, t->crs_ext.data(), ell + 9, bits, &t->table);
    if (rc != CDP_OK) {
        cdp_ctx_destroy(t->ctx);
        delete t;
        return rc;
    }
    crs_table_list().push_back(t);
    *out = t;
    return CDP_OK;
}
inline void crs_table_release(SharedCrsTable *t) {
    if (!t) return;
    std::lock_guard<std::mutex> lock(crs_table_mutex());
    if (--t->refs > 0) return;
    auto &v = crs_table_list();
    for (size_t i = 0; i < v.size(); i++)
        if (v[i] == t) { v.erase(v.begin() + i); break; }
    cdp_fixed_table_destroy(t->ctx, t->table);
    cdp_ctx_destroy(t->ctx);
    delete t;
}

}  // namespace cdp_host
