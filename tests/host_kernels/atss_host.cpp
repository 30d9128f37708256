// Test infrastructure: serial HOST driver around super_gradients_b200/csrc/atss_math.cuh (the arithmetic of the CUDA ATSS
// assigner), compiled with g++ by tests/host_atss.py.  Same steps as atss_candidates_kernel / atss_resolve_kernel.
#include <cmath>
#include <limits>
#include <vector>

#include "sgb200.h"
#include "atss_math.cuh"

using sgb_atss::Box;

extern "C" int atss_assign_host(const SgbLossDesc* d, const float* reg, const float* anchors, const float* ap, const float* st,
                                const int32_t* level_sizes, int32_t n_levels, const float* gtb, const int32_t* gtl, const uint8_t* gtv,
                                int32_t* alabel, float* abox, float* ascore, double* sums) {
  const int64_t BL = (int64_t)d->B * d->L;
  std::vector<int> count(BL, 0), owner(BL, 0x7fffffff);
  if (n_levels > sgb_atss::kMaxLevels || d->topk > sgb_atss::kMaxTopk) return 1;
  std::vector<int> start(n_levels + 1, 0);
  for (int i = 0; i < n_levels; ++i) {
    if (level_sizes[i] < d->topk) return 2;
    start[i + 1] = start[i] + level_sizes[i];
  }
  if (start[n_levels] != d->L) return 3;
  if (d->n_max == 0) {
    for (int64_t i = 0; i < BL; ++i) {
      alabel[i] = d->ncls;
      ascore[i] = 0.f;
      for (int k = 0; k < 4; ++k) abox[i * 4 + k] = 0.f;
    }
    return 0;
  }
  for (int bg = 0; bg < d->B * d->n_max; ++bg) {
    if (!gtv[bg]) continue;
    const int b = bg / d->n_max, g = bg - b * d->n_max;
    const Box gt = sgb_atss::load_box(gtb + (int64_t)bg * 4);
    std::vector<int> cand;
    for (int lvl = 0; lvl < n_levels; ++lvl) {
      const int base = start[lvl], num = start[lvl + 1] - base;
      std::vector<float> dist(num);
      for (int a = 0; a < num; ++a) dist[a] = sgb_atss::center_distance(gt, sgb_atss::load_box(anchors + (int64_t)(base + a) * 4));
      for (int k = 0; k < d->topk; ++k) {
        float bv = std::numeric_limits<float>::infinity();
        int bi = 0x7fffffff;
        for (int a = 0; a < num; ++a)
          if (dist[a] < bv) {
            bv = dist[a];
            bi = a;
          }
        cand.push_back(base + bi);
        dist[bi] = std::numeric_limits<float>::infinity();
      }
    }
    const int K = (int)cand.size();
    std::vector<float> ciou(K);
    for (int t = 0; t < K; ++t) ciou[t] = sgb_atss::iou(gt, sgb_atss::load_box(anchors + (int64_t)cand[t] * 4), 1e-10f);
    const float thr = sgb_atss::iou_threshold(ciou.data(), K);
    for (int t = 0; t < K; ++t)
      if (ciou[t] > thr && sgb_atss::center_inside(sgb_atss::load_box(anchors + (int64_t)cand[t] * 4), gt)) {
        const int64_t i = (int64_t)b * d->L + cand[t];
        count[i] += 1;
        owner[i] = g < owner[i] ? g : owner[i];
      }
  }
  const int bins = d->reg_max + 1;
  double total = 0.0;
  for (int64_t i = 0; i < BL; ++i) {
    const int l = (int)(i % d->L), b = (int)(i / d->L);
    const float* gts = gtb + (int64_t)b * d->n_max * 4;
    int ag = -1;
    if (count[i] == 1) ag = owner[i];
    else if (count[i] > 1) ag = sgb_atss::argmax_iou_gt(sgb_atss::load_box(anchors + (int64_t)l * 4), gts, d->n_max);
    const float* gb = gts + (ag >= 0 ? ag : 0) * 4;
    for (int k = 0; k < 4; ++k) abox[i * 4 + k] = gb[k];
    float sc = 0.f;
    int lab = d->ncls;
    if (ag >= 0) {
      lab = gtl[b * d->n_max + ag];
      const Box p = sgb_atss::decode_box(reg + i * 4 * bins, bins, ap[l * 2], ap[l * 2 + 1], st[l]);
      sc = sgb_atss::iou(sgb_atss::load_box(gb), p, 1e-9f);
    }
    alabel[i] = lab;
    ascore[i] = sc;
    total += sc;
  }
  sums[3] += total;
  return 0;
}
