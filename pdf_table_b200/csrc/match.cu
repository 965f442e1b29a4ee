// Cell / text matching: the first post-model stage of the reference's table export,
// OcrTableToHtmlTask.find_top1_mach_box (ocr_pdf/ocr_table_to_html_task.py:48-77) for every recognised text box of a table:
//   * the FIRST table cell (in list order) that contains the text box -- box_in_other_box(cell, text, diff = 2)
//     (pdf_table/table_common.py:138-160) -- wins outright;
//   * otherwise the cell with the smallest key (1 - compute_iou_v2(text, cell), distance(text, cell)) in lexicographic order,
//     first occurrence on ties (sorted() is stable and list.index returns the first equal tuple; table_common.py:435-441, 473-516).
// One warp per text box, lanes stride over the cells, all arithmetic in float64 with round-to-nearest intrinsics in the
// reference's operation order (Python floats), so the chosen indices are identical; the key is reduced with a warp shuffle.
#include "engine.h"

namespace dv {
namespace {

struct MatchKey {
    double a, b;  // (1 - iou, distance)
    int idx;
};

__device__ __forceinline__ bool key_less(const MatchKey& x, const MatchKey& y) {
    if (x.a != y.a) return x.a < y.a;
    if (x.b != y.b) return x.b < y.b;
    return x.idx < y.idx;
}

__global__ void __launch_bounds__(128)
k_match_cells(const double* __restrict__ text, int n_text, const double* __restrict__ cells, int n_cells, int32_t* __restrict__ out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_text) return;
    const double x3 = text[4 * warp], y3 = text[4 * warp + 1], x4 = text[4 * warp + 2], y4 = text[4 * warp + 3];
    const double min_y2 = fmin(y3, y4), max_y2 = fmax(y3, y4);
    const double t_area = fabs(__dmul_rn(__dsub_rn(x4, x3), __dsub_rn(y4, y3)));
    int first_in = 0x7fffffff;
    MatchKey best{1e300, 1e300, 0x7fffffff};
    for (int c = lane; c < n_cells; c += 32) {
        const double x1 = cells[4 * c], y1 = cells[4 * c + 1], x2 = cells[4 * c + 2], y2 = cells[4 * c + 3];
        // box_in_other_box(cell, text): x3 >= x1 - 2 and x4 <= x2 + 2 and min_y_1 - 2 <= min_y_2 <= max_y_2 <= max_y_1 + 2
        const double min_y1 = fmin(y1, y2), max_y1 = fmax(y1, y2);
        if (x3 >= __dsub_rn(x1, 2.0) && x4 <= __dadd_rn(x2, 2.0) && __dsub_rn(min_y1, 2.0) <= min_y2 && min_y2 <= max_y2 &&
            max_y2 <= __dadd_rn(max_y1, 2.0)) {
            first_in = min(first_in, c);
            continue;
        }
        // distance(text, cell): box_1 = text, box_2 = cell
        const double ax = fabs(__dsub_rn(x1, x3)), ay = fabs(__dsub_rn(y1, y3)), bx = fabs(__dsub_rn(x2, x4)), by = fabs(__dsub_rn(y2, y4));
        const double dis = __dadd_rn(__dadd_rn(__dadd_rn(ax, ay), bx), by);
        const double dist = __dadd_rn(dis, fmin(__dadd_rn(ax, ay), __dadd_rn(bx, by)));
        // compute_iou_v2(text, cell)
        const double ix1 = fmax(x3, x1), iy1 = fmax(y3, y1), ix2 = fmin(x4, x2), iy2 = fmin(y4, y2);
        double dx = __dsub_rn(ix2, ix1), dy = __dsub_rn(iy2, iy1);
        if (dx < 0) dx = 0;
        if (dy < 0) dy = 0;
        const double inter = __dmul_rn(dx, dy);
        const double c_area = fabs(__dmul_rn(__dsub_rn(x2, x1), __dsub_rn(y2, y1)));
        const double iou = __ddiv_rn(inter, __dadd_rn(__dsub_rn(__dadd_rn(t_area, c_area), inter), 1e-6));
        const MatchKey k{__dsub_rn(1.0, iou), dist, c};
        if (key_less(k, best)) best = k;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        first_in = min(first_in, __shfl_xor_sync(0xffffffffu, first_in, o));
        MatchKey other{__shfl_xor_sync(0xffffffffu, best.a, o), __shfl_xor_sync(0xffffffffu, best.b, o), __shfl_xor_sync(0xffffffffu, best.idx, o)};
        if (key_less(other, best)) best = other;
    }
    // the reference scans the cells in order and stops at the first containing one; cells before it only matter if none contains
    if (lane == 0) out[warp] = first_in != 0x7fffffff ? first_in : (best.idx != 0x7fffffff ? best.idx : -1);
}

}  // namespace

int match_cells(Engine* e, const double* text_boxes, int n_text, const double* cell_boxes, int n_cells, int32_t* top1_out) {
    if (n_text == 0) return 0;
    if (!text_boxes || !cell_boxes || !top1_out || n_text < 0 || n_cells <= 0) return set_err(e, DV_ERR_ARG, "match_cells: bad arguments");
    e->launch_begin("k_match_cells", "match", 0.0, 32.0 * n_text + 32.0 * n_cells + 4.0 * n_text);
    k_match_cells<<<(n_text * 32 + 127) / 128, 128, 0, e->stream>>>(text_boxes, n_text, cell_boxes, n_cells, top1_out);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

}  // namespace dv
