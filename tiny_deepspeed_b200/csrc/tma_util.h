// Host helper shared by the TMA-fed kernels: 4-D bf16 tensor map (inner, rows, inner-batch, outer-batch), SWIZZLE_128B.
#pragma once
#include <cuda.h>

#include "kernels.h"

namespace tds {

// K-major operand (stored [rows][K]):  dims (K, rows, nb2, nb1), box (64, box_rows_kmajor, 1, 1)
// MN-major operand (stored [K][rows]): dims (rows, K, nb2, nb1), box (64, box_krows_mnmajor, 1, 1)
bool make_map(CUtensorMap* out, const GemmOperand& op, int rows_mn, int K, int nb1, int nb2, int box_rows_kmajor,
              int box_krows_mnmajor = 64, bool f32 = false);
// With f32 = true the operand holds fp32 (fed to kind::tf32 MMAs): a 128-byte swizzle row is 32 elements, so the inner box
// extent is 32 in both majors and byte strides scale by 4.

// 2-D fp32 map [rows][cols] (row pitch `ld` elements), SWIZZLE_128B, box (box_cols <= 32, box_rows)
bool make_map_f32_2d(CUtensorMap* out, void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows);

}  // namespace tds
