// match.cuh -- sdnet_match_kernel: the reference evaluator's nearest-ground-truth matching on the packed detections.
#pragma once

namespace {

// ---------------------------------------------------------------------------------------------
// evaluator matching (the step after the path): reference Evaluator.eval_anchor / eval_part,
// src/sdnet/model/evaluator.py:244-334, on the packed detections.  One CTA per image; the greedy
// "first detection in score order that claims a ground truth wins it" loop is order-free once every
// detection knows its nearest ground truth: the winner of ground truth j is the smallest slot index
// among the detections whose nearest is j and whose distance is under the threshold (atomicMin).
// Arithmetic is double throughout, like the reference's Python floats.
// ---------------------------------------------------------------------------------------------
constexpr int kMatchThreads = 256;

__device__ void match_pass(const float* __restrict__ out, int row_len, int slots, int n_classes, bool strict_gt,
                           double conf, double sx, double sy, const double* __restrict__ scale,
                           const double* __restrict__ gt, int n_gt, int* __restrict__ stats_out,
                           double* __restrict__ acc_out, double* s_gx, double* s_gy, int* s_glab, int* s_winner,
                           int* s_jmin, double* s_dist, int* s_stats) {
  const double rx = scale[0], ry = scale[1], thresh = scale[2], norm = scale[3];
  for (int j = threadIdx.x; j < n_gt; j += blockDim.x) {
    s_gx[j] = gt[3 * j + 0] * rx;  // annotation.resized(...): evaluator.py:247
    s_gy[j] = gt[3 * j + 1] * ry;
    s_glab[j] = (int)gt[3 * j + 2];
    s_winner[j] = 0x7fffffff;
  }
  for (int i = threadIdx.x; i < 3 * n_classes; i += blockDim.x) s_stats[i] = 0;
  __syncthreads();
  for (int j = threadIdx.x; j < n_gt; j += blockDim.x)
    if (s_glab[j] >= 0 && s_glab[j] < n_classes) atomicAdd(&s_stats[3 * s_glab[j] + 1], 1);  // npos
  for (int i = threadIdx.x; i < slots; i += blockDim.x) {
    const float* row = out + (size_t)i * row_len;
    const double score = (double)row[2];
    const bool det = strict_gt ? (score > conf) : !(score < conf);
    const int lab = (int)row[3];
    int jmin = -1;
    double best = 1.7976931348623157e308;  // sys.float_info.max
    if (det) {
      atomicAdd(&s_stats[3 * lab + 0], 1);  // ndet
      const double px = ((double)row[0] * sx) * rx, py = ((double)row[1] * sy) * ry;  // decoders.py:139 then evaluator.py:248
      for (int j = 0; j < n_gt; ++j) {
        if (s_glab[j] != lab) continue;
        const double d = hypot(px - s_gx[j], py - s_gy[j]);  // np.hypot, utils.py:31-32
        if (d < best) { best = d; jmin = j; }
      }
      if (jmin >= 0 && best < thresh) atomicMin(&s_winner[jmin], i);
    }
    s_jmin[i] = det ? jmin : -1;
    s_dist[i] = best;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < slots; i += blockDim.x) {
    const int jmin = s_jmin[i];
    const bool tp = jmin >= 0 && s_dist[i] < thresh && s_winner[jmin] == i;
    acc_out[i] = tp ? s_dist[i] / norm : __longlong_as_double(0x7ff8000000000000ll);
    if (tp) atomicAdd(&s_stats[3 * (int)out[(size_t)i * row_len + 3] + 2], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * n_classes; i += blockDim.x) stats_out[i] = s_stats[i];
  __syncthreads();
}

__global__ void __launch_bounds__(kMatchThreads) sdnet_match_kernel(const SdnetMatchParams p) {
  __shared__ double s_gx[SDNET_MAX_GT], s_gy[SDNET_MAX_GT], s_dist[SDNET_MAX_TOPK];
  __shared__ int s_glab[SDNET_MAX_GT], s_winner[SDNET_MAX_GT], s_jmin[SDNET_MAX_TOPK], s_stats[3 * SDNET_MAX_CHANNELS];
  const int b = blockIdx.x;
  const double* scale = p.image_scale + 4 * (size_t)b;
  match_pass(p.anchor_out + (size_t)b * p.K * 4, 4, p.K, p.M, true, p.conf, p.sx, p.sy, scale,
             p.gt_anchors + (size_t)b * p.max_gt_anchors * 3, min(p.n_gt_anchors[b], p.max_gt_anchors),
             p.anchor_stats + (size_t)b * p.M * 3, p.anchor_acc + (size_t)b * p.K, s_gx, s_gy, s_glab, s_winner, s_jmin,
             s_dist, s_stats);
  match_pass(p.part_out + (size_t)b * p.P * 6, 6, p.P, p.N, false, p.conf, p.sx, p.sy, scale,
             p.gt_parts + (size_t)b * p.max_gt_parts * 3, min(p.n_gt_parts[b], p.max_gt_parts),
             p.part_stats + (size_t)b * p.N * 3, p.part_acc + (size_t)b * p.P, s_gx, s_gy, s_glab, s_winner, s_jmin, s_dist,
             s_stats);
}


}  // namespace
