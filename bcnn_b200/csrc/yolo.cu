// yolo.cu -- YOLOv3 detection loss on the device (sm_100a).
//
// The reference computes this loss on the host even in its CUDA build: it copies the activated
// head to the host, runs the loops of bcnn_forward_yolo_layer_cpu (src/layers/bcnn_yolo.c:251-416)
// and copies the gradient back (:418-431) -- a device-to-host synchronisation in the middle of
// every training step. Here the same two sweeps are kernels on the net's stream:
//   yolo_loss_cells_kernel   one thread per (sample, anchor, cell): zero the cell's gradient
//                            entries, decode its box, best IoU against the sample's truth list,
//                            objectness gradient (0 when some truth overlaps by more than 0.5);
//   yolo_loss_truths_kernel  one warp per sample walks that sample's truths IN ORDER (later boxes
//                            overwrite earlier ones, as on the host): best anchor by size, and when
//                            it belongs to this head the box / objectness / class targets of the
//                            cell the box centre falls in;
//   yolo_cost_*_kernel       cost = sum of squared gradient entries, per-CTA partials folded in a
//                            fixed order (deterministic).
// Arithmetic: the reference's float operations in its order, with explicit round-to-nearest
// multiplies / adds / divides so nothing is contracted into FMAs; expf / logf are CUDA's (<= 2 ulp
// from glibc's), so gradients match the host restatement (bcnn_yolo_loss_host, bit-identical to the
// reference) to ~1e-6 rather than bit for bit. Algorithmic traffic: 12 B per head element (read the
// head, write the gradient, read it again for the cost).
#include "common.cuh"

using namespace b200;

namespace {

struct Box { float x, y, w, h; };

__device__ __forceinline__ float span_overlap(float c1, float e1, float c2, float e2) {
    const float h1 = __fdiv_rn(e1, 2.f), h2 = __fdiv_rn(e2, 2.f);
    const float lo1 = __fsub_rn(c1, h1), lo2 = __fsub_rn(c2, h2);
    const float hi1 = __fadd_rn(c1, h1), hi2 = __fadd_rn(c2, h2);
    return __fsub_rn(hi1 < hi2 ? hi1 : hi2, lo1 > lo2 ? lo1 : lo2);
}

__device__ __forceinline__ float box_iou(Box a, Box b) {
    const float w = span_overlap(a.x, a.w, b.x, b.w), h = span_overlap(a.y, a.h, b.y, b.h);
    const float inter = (w < 0 || h < 0) ? 0.f : __fmul_rn(w, h);
    const float uni = __fsub_rn(__fadd_rn(__fmul_rn(a.w, a.h), __fmul_rn(b.w, b.h)), inter);
    return __fdiv_rn(inter, uni);
}

// get_yolo_box (reference :137-146): `e` points at the cell's first entry, entries hw apart.
__device__ __forceinline__ Box decode_box(const float *e, int hw, float aw, float ah, int col,
                                          int row, int lw, int lh, int netw, int neth) {
    Box b;
    b.x = __fdiv_rn(__fadd_rn((float)col, e[0]), (float)lw);
    b.y = __fdiv_rn(__fadd_rn((float)row, e[hw]), (float)lh);
    b.w = __fdiv_rn(__fmul_rn(expf(e[2 * (size_t)hw]), aw), (float)netw);
    b.h = __fdiv_rn(__fmul_rn(expf(e[3 * (size_t)hw]), ah), (float)neth);
    return b;
}

struct YoloDims {
    int n, num, classes, coords, lw, lh, netw, neth, total, max_boxes, truths;
};

__global__ void __launch_bounds__(256)
yolo_loss_cells_kernel(const float *__restrict__ out, const float *__restrict__ label,
                       const float *__restrict__ anchors, const int *__restrict__ mask,
                       float *__restrict__ delta, YoloDims d) {
    const int hw = d.lw * d.lh, group = d.coords + d.classes + 1;
    const size_t cells = (size_t)d.n * d.num * hw;
    for (size_t t = (size_t)blockIdx.x * 256 + threadIdx.x; t < cells;
         t += (size_t)gridDim.x * 256) {
        const int pos = (int)(t % hw);
        const int a = (int)((t / hw) % d.num);
        const int b = (int)(t / ((size_t)hw * d.num));
        const size_t cell = ((size_t)b * d.num + a) * group * hw + pos;
        for (int e = 0; e < group; ++e) delta[cell + (size_t)e * hw] = 0.f;
        const int anchor = mask[a];
        const Box pred = decode_box(out + cell, hw, anchors[2 * anchor], anchors[2 * anchor + 1],
                                    pos % d.lw, pos / d.lw, d.lw, d.lh, d.netw, d.neth);
        const float *truths = label + (size_t)b * d.truths;
        float best_iou = 0.f;
        for (int k = 0; k < d.max_boxes; ++k) {
            const float *f = truths + (size_t)k * (d.coords + 1);
            if (f[0] == 0.f) break;
            const Box truth = {f[0], f[1], f[2], f[3]};
            const float iou = box_iou(pred, truth);
            if (iou > best_iou) best_iou = iou;
        }
        const size_t obj = cell + (size_t)d.coords * hw;
        delta[obj] = best_iou > 0.5f ? 0.f : out[obj];
    }
}

// One warp per sample; lane 0 does the scalar work, all lanes share the class loop.
__global__ void __launch_bounds__(32)
yolo_loss_truths_kernel(const float *__restrict__ out, const float *__restrict__ label,
                        const float *__restrict__ anchors, const int *__restrict__ mask,
                        float *delta, YoloDims d) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const int hw = d.lw * d.lh, group = d.coords + d.classes + 1;
    const float *truths = label + (size_t)b * d.truths;
    const size_t base = (size_t)b * d.num * group * hw;
    for (int k = 0; k < d.max_boxes; ++k) {
        const float *f = truths + (size_t)k * (d.coords + 1);
        if (f[0] == 0.f) break;
        const Box truth = {f[0], f[1], f[2], f[3]};
        const int col = (int)__fmul_rn(truth.x, (float)d.lw), row = (int)__fmul_rn(truth.y, (float)d.lh);
        if (col < 0 || col >= d.lw || row < 0 || row >= d.lh) continue;
        const Box centred = {0.f, 0.f, truth.w, truth.h};
        float best_iou = 0.f;
        int best = 0;
        for (int n = 0; n < d.total; ++n) {
            const Box prior = {0.f, 0.f, __fdiv_rn(anchors[2 * n], (float)d.netw),
                               __fdiv_rn(anchors[2 * n + 1], (float)d.neth)};
            const float iou = box_iou(prior, centred);
            if (iou > best_iou) { best_iou = iou; best = n; }
        }
        int a = -1;
        for (int m = 0; m < d.num && a < 0; ++m)
            if (mask[m] == best) a = m;
        if (a < 0) continue;  // another head owns this anchor
        const size_t cell = base + (size_t)a * group * hw + (size_t)row * d.lw + col;
        const size_t obj = cell + (size_t)d.coords * hw, first_class = obj + hw;
        const int cls = (int)f[d.coords];
        const bool claimed = delta[first_class] != 0.f;  // read by every lane before any write
        __syncwarp();
        if (lane == 0) {
            const float aw = anchors[2 * best], ah = anchors[2 * best + 1];
            const float scale = __fsub_rn(2.f, __fmul_rn(truth.w, truth.h));
            const float target[4] = {
                __fsub_rn(__fmul_rn(truth.x, (float)d.lw), (float)col),
                __fsub_rn(__fmul_rn(truth.y, (float)d.lh), (float)row),
                logf(__fdiv_rn(__fmul_rn(truth.w, (float)d.netw), aw)),
                logf(__fdiv_rn(__fmul_rn(truth.h, (float)d.neth), ah))};
            for (int e = 0; e < 4; ++e)
                delta[cell + (size_t)e * hw] =
                    __fmul_rn(-scale, __fsub_rn(target[e], out[cell + (size_t)e * hw]));
            delta[obj] = __fsub_rn(out[obj], 1.f);
        }
        if (claimed) {  // cell already claimed by an earlier truth: only this class is pushed up
            if (lane == 0 && cls >= 0 && cls < d.classes)
                delta[first_class + (size_t)hw * cls] = __fsub_rn(out[first_class + (size_t)hw * cls], 1.f);
        } else {
            for (int n = lane; n < d.classes; n += 32)
                delta[first_class + (size_t)hw * n] =
                    __fsub_rn(out[first_class + (size_t)hw * n], n == cls ? 1.f : 0.f);
        }
        __syncwarp();  // this truth's stores are visible to the next truth's `claimed` read
    }
}

__global__ void __launch_bounds__(256)
yolo_cost_partial_kernel(const float *__restrict__ delta, size_t total, float *__restrict__ partial) {
    __shared__ float red[8];
    float acc[1] = {0.f};
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256)
        acc[0] = fmaf(delta[i], delta[i], acc[0]);
    block_sum<1, 256>(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc[0];
}

__global__ void yolo_cost_final_kernel(const float *__restrict__ partial, int count,
                                       float *__restrict__ cost) {
    float s = 0.f;
    for (int i = 0; i < count; ++i) s += partial[i];  // fixed order
    cost[0] = s;
}

}  // namespace

extern "C" int bcnn_b200_yolo_cost_scratch_floats(void) { return 1 + 148; }

extern "C" int bcnn_b200_yolo_loss_forward(const float *out, const float *label, const float *anchors,
                                           const int *mask, float *delta, float *cost_scratch,
                                           int n, int boxes_per_cell, int classes, int coords,
                                           int lw, int lh, int netw, int neth, int total_anchors,
                                           int max_boxes, void *stream) {
    const YoloDims d = {n, boxes_per_cell, classes, coords, lw, lh, netw, neth, total_anchors,
                        max_boxes, max_boxes * (coords + 1)};
    const size_t cells = (size_t)n * boxes_per_cell * lw * lh;
    const size_t total = cells * (size_t)(coords + classes + 1);
    if (total == 0) return 0;
    if (coords < 4 || total >= (1ull << 32)) return (int)cudaErrorInvalidValue;
    cudaStream_t s = as_stream(stream);
    yolo_loss_cells_kernel<<<stream_grid(cells, 256), 256, 0, s>>>(out, label, anchors, mask, delta, d);
    ++g_launch_count;
    yolo_loss_truths_kernel<<<n, 32, 0, s>>>(out, label, anchors, mask, delta, d);
    ++g_launch_count;
    const int blocks = 148;
    yolo_cost_partial_kernel<<<blocks, 256, 0, s>>>(delta, total, cost_scratch + 1);
    ++g_launch_count;
    yolo_cost_final_kernel<<<1, 1, 0, s>>>(cost_scratch + 1, blocks, cost_scratch);
    return launched();
}
