// Host-side helpers shared by the C-ABI translation units: error reporting, TMA descriptor encoding
// (driver entry point resolved at run time so the library links against cudart only), device queries.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/tokensgen_b200.h"

namespace tg {

// Sets the thread-local message returned by tg_last_error(); returns `code` so callers can `return fail(...)`.
int fail(int code, const char* fmt, ...);
// Checks the launch of the kernel just enqueued (cudaGetLastError); 0 when fine.
int check_launch(const char* what);

int sm_count();

// tg_rowmap sequence-parallel shard (row0, rows_local): rows_local == 0 means "all rows of the batch".
inline bool rowmap_shard_ok(const tg_rowmap* m) {
    if (m->rows_local == 0) return m->row0 == 0;
    return m->row0 >= 0 && m->rows_local > 0 && m->row0 + m->rows_local <= m->rows_per_batch;
}
inline tg_rowmap normalised_rowmap(const tg_rowmap* m) {
    tg_rowmap r = *m;
    if (r.rows_local == 0) { r.rows_local = r.rows_per_batch; r.row0 = 0; }
    return r;
}

// Row-major bf16 tensor maps with 128-byte swizzle and zero fill of out-of-bounds elements.
//   2D: dims (inner, rows), box (box_inner, box_rows), row stride in bytes.
//   3D: dims (inner, rows, batch), box (box_inner, box_rows, 1).
// Returns 0 or a positive CUresult (message already set).
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                 uint32_t box_inner, uint32_t box_rows);
int make_tmap_3d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t batch,
                 uint64_t row_stride_bytes, uint64_t batch_stride_bytes, uint32_t box_inner, uint32_t box_rows);

// General form (rank <= 5): dims / box / element strides innermost first, strides_bytes for dims 1..rank-1.
int make_tmap_nd(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                 const uint32_t* box, const uint32_t* elem_strides);

}  // namespace tg
