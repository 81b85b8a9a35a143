// Column copies between rollout-shaped tensors: the environment migration of the balanced sharded recurrent update
// (SURVEY.md section 8e).  A rollout field is [rows, N, w bytes] (time-major, reference storage.py:10-30); a migration
// record is the same data environment-major ([n, rows, w], one record per travelling environment).  Packing records,
// unpacking them into the guest columns of the shadow rollout and copying freshly uploaded columns into it are all
//     dst[dst_col(j)][t] = src[src_col(j)][t]        j < n_cols, t < rows,
// with byte strides per column and per row on both sides and optional column lists -- ONE launch for all fields of a
// rollout instead of four torch kernels (index_select, transpose copy, reshape copy, strided copy) per field.
#include "common.cuh"
#include "decls.cuh"

namespace a2cb {

namespace {

constexpr int kMaxFields = 12;
constexpr int kThreads = 256;

struct ColCopyArgs {
    ColCopyField f[kMaxFields];
    long first_unit[kMaxFields + 1];       // prefix sums of n_cols * rows over the fields
    int n_fields;
};

__global__ void __launch_bounds__(kThreads) colcopy_kernel(ColCopyArgs a, const long* __restrict__ dst_cols,
                                                           const long* __restrict__ src_cols, int n_cols) {
    const long total = a.first_unit[a.n_fields];
    for (long u = blockIdx.x; u < total; u += gridDim.x) {
        int fi = 0;
        while (u >= a.first_unit[fi + 1]) ++fi;
        const ColCopyField f = a.f[fi];
        const long v = u - a.first_unit[fi];
        if (f.wide) {                                          // one unit = one (column, row) element of w bytes, w % 16 == 0
            const long j = v / f.rows, t = v - j * f.rows;
            const long dc = dst_cols ? dst_cols[j] : j, sc = src_cols ? src_cols[j] : j;
            const int4* s = reinterpret_cast<const int4*>(f.src + sc * f.src_cs + t * f.src_rs);
            int4* d = reinterpret_cast<int4*>(f.dst + dc * f.dst_cs + t * f.dst_rs);
            const int n16 = (int)(f.w >> 4);
            for (int i = threadIdx.x; i < n16; i += kThreads) d[i] = s[i];
        } else {                                               // narrow: one unit = kThreads elements, one thread each
            const long e = v * kThreads + threadIdx.x;
            if (e < (long)n_cols * f.rows) {
                const long j = e / f.rows, t = e - j * f.rows;
                const long dc = dst_cols ? dst_cols[j] : j, sc = src_cols ? src_cols[j] : j;
                const char* s = f.src + sc * f.src_cs + t * f.src_rs;
                char* d = f.dst + dc * f.dst_cs + t * f.dst_rs;
                if (f.w == 4) *reinterpret_cast<int*>(d) = *reinterpret_cast<const int*>(s);
                else if (f.w == 8) *reinterpret_cast<long*>(d) = *reinterpret_cast<const long*>(s);
                else for (long b = 0; b < f.w; ++b) d[b] = s[b];
            }
        }
    }
}

}  // namespace

int copy_columns(const ColCopyField* fields, int n_fields, const long* dst_cols, const long* src_cols, int n_cols,
                 cudaStream_t st) {
    A2CB_REQUIRE(fields && n_fields >= 1 && n_fields <= kMaxFields, "copy_columns: 1 .. %d fields", kMaxFields);
    A2CB_REQUIRE(n_cols >= 0, "copy_columns: negative column count");
    if (n_cols == 0) return A2CB_OK;
    ColCopyArgs a;
    a.n_fields = n_fields;
    a.first_unit[0] = 0;
    for (int i = 0; i < n_fields; ++i) {
        ColCopyField f = fields[i];
        A2CB_REQUIRE(f.dst && f.src && f.w > 0 && f.rows > 0, "copy_columns: field %d: null pointer or empty element", i);
        const bool al = !(((uintptr_t)f.dst | (uintptr_t)f.src | (uintptr_t)f.dst_cs | (uintptr_t)f.dst_rs | (uintptr_t)f.src_cs |
                           (uintptr_t)f.src_rs | (uintptr_t)f.w) & 15);
        f.wide = (al && f.w >= 256) ? 1 : 0;
        if (!f.wide) {
            const bool a4 = !(((uintptr_t)f.dst | (uintptr_t)f.src | (uintptr_t)f.dst_cs | (uintptr_t)f.dst_rs | (uintptr_t)f.src_cs |
                               (uintptr_t)f.src_rs) & (f.w == 8 ? 7 : 3));
            A2CB_REQUIRE((f.w != 4 && f.w != 8) || a4, "copy_columns: field %d: 4 / 8-byte elements must be naturally aligned", i);
            A2CB_REQUIRE(f.w <= 4096, "copy_columns: field %d: %ld-byte elements need 16-byte aligned strides", i, (long)f.w);
        }
        a.f[i] = f;
        const long elems = (long)n_cols * f.rows;
        a.first_unit[i + 1] = a.first_unit[i] + (f.wide ? elems : (elems + kThreads - 1) / kThreads);
    }
    const long total = a.first_unit[n_fields];
    const long cap = (long)kNumSM * 16;
    colcopy_kernel<<<(unsigned)(total < cap ? total : cap), kThreads, 0, st>>>(a, dst_cols, src_cols, n_cols);
    return check_launch("colcopy");
}

}  // namespace a2cb
