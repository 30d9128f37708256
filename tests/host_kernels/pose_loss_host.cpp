// Test infrastructure: a serial HOST driver around super_gradients_b200/csrc/pose_loss_math.cuh, the header that holds the
// per-pair / per-anchor arithmetic of the CUDA pose-loss kernels (pose_loss.cu).  Compiled with g++ by
// tests/test_pose_loss_host.py and compared with the oracle and the reference-generated fixtures, so the kernels'
// arithmetic (not their parallel schedule) is checked on machines without a GPU.  Same stages as the kernels: decode ->
// per-gt top-k (ties: lowest anchor index) -> per-anchor resolve -> finish -> loss + gradients -> finalize.
#include <cstring>
#include <vector>

#include "pose_loss_math.cuh"

using namespace sgb_pose;

extern "C" int pose_loss_host(const SgbPoseLossDesc* dp, const float* cls, const float* reg, const float* pose, const float* plog,
                              const float* ap, const float* st, const float* gtb, const float* gtp, const uint8_t* gtc,
                              const uint8_t* gtv, const float* sigmas, float grad_scale, int32_t* assigned_gt,
                              float* assigned_score, double* sums, float* gcls, float* greg, float* gpose, float* gplog,
                              float* items) {
  const SgbPoseLossDesc d = *dp;
  const int nb = d.reg_max + 1;
  const int64_t BL = (int64_t)d.B * d.L;
  for (int k = 0; k < 8; ++k) sums[k] = 0.0;
  std::memset(greg, 0, sizeof(float) * BL * 4 * nb);
  std::memset(gpose, 0, sizeof(float) * BL * d.J * 2);
  std::memset(gplog, 0, sizeof(float) * BL * d.J);
  if (d.n_max == 0) {
    for (int64_t i = 0; i < BL; ++i) {
      assigned_gt[i] = -1;
      assigned_score[i] = 0.f;
    }
  } else {
    std::vector<float> pbox(BL * 4), apair(BL * 2), gmax((size_t)d.B * d.n_max * 2, 0.f), smet(d.L);
    std::vector<int> topk((size_t)d.B * d.n_max * d.topk, -1), agt(BL);
    for (int64_t i = 0; i < BL; ++i) {
      const int l = i % d.L;
      decode_box(reg + i * 4 * nb, nb, ap[l * 2], ap[l * 2 + 1], st[l], &pbox[i * 4]);
    }
    for (int bg = 0; bg < d.B * d.n_max; ++bg) {
      if (!gtv[bg]) continue;
      const int b = bg / d.n_max;
      const PBox g{gtb[bg * 4 + 0], gtb[bg * 4 + 1], gtb[bg * 4 + 2], gtb[bg * 4 + 3]};
      for (int l = 0; l < d.L; ++l) {
        const int64_t i = (int64_t)b * d.L + l;
        const PBox p{pbox[i * 4 + 0], pbox[i * 4 + 1], pbox[i * 4 + 2], pbox[i * 4 + 3]};
        const float iou = pair_iou(d, g, gtp + (int64_t)bg * d.J * 3, p, pose + i * d.J * 2, sigmas);
        smet[l] = tal_metric(d, sigmoid_f(cls[i]), iou) * (inside_gt(ap[l * 2], ap[l * 2 + 1], g) ? 1.f : 0.f);
      }
      for (int k = 0; k < d.topk; ++k) {
        float bv = -1.f;
        int bi = 0x7fffffff;
        for (int l = 0; l < d.L; ++l)
          if (smet[l] > bv) {
            bv = smet[l];
            bi = l;
          }
        topk[(size_t)bg * d.topk + k] = bi;
        smet[bi] = -2.f;
      }
    }
    for (int64_t i = 0; i < BL; ++i) {
      const int l = i % d.L, b = i / d.L;
      int ag;
      float met, iou;
      resolve_anchor(d, b, l, pbox.data(), cls, pose, ap, gtb, gtp, gtv, sigmas, topk.data(), &ag, &met, &iou);
      agt[i] = ag;
      apair[i * 2] = met;
      apair[i * 2 + 1] = iou;
      if (ag >= 0) {
        float* gm = &gmax[((size_t)b * d.n_max + ag) * 2];
        gm[0] = fmaxf(gm[0], met);
        gm[1] = fmaxf(gm[1], iou);
      }
    }
    for (int64_t i = 0; i < BL; ++i) {
      const int b = i / d.L, ag = agt[i];
      const size_t bg = (size_t)b * d.n_max + (ag >= 0 ? ag : 0);
      int pos;
      float sc;
      finish_anchor(ag, apair[i * 2], gmax[bg * 2], gmax[bg * 2 + 1], ag >= 0 && gtc[bg] != 0, &pos, &sc);
      assigned_gt[i] = pos;
      assigned_score[i] = sc;
      sums[3] += sc;
      sums[6] += pos >= 0 ? 1.0 : 0.0;
    }
  }
  const double nrm = sums[3] < 1.0 ? 1.0 : sums[3], npos = sums[6] < 1.0 ? 1.0 : sums[6];
  const float inv_norm = grad_scale / (float)nrm, inv_pos = grad_scale / (float)npos;
  for (int64_t i = 0; i < BL; ++i) {
    AnchorSums acc{0.f, 0.f, 0.f, 0.f, 0.f};
    anchor_loss(d, (int)(i / d.L), (int)(i % d.L), cls, reg, pose, plog, ap, st, gtb, gtp, sigmas, assigned_gt[i], assigned_score[i],
                inv_norm, inv_pos, gcls, greg, gpose, gplog, &acc);
    sums[0] += acc.cls;
    sums[1] += acc.iou;
    sums[2] += acc.dfl;
    sums[4] += acc.pcls;
    sums[5] += acc.preg;
  }
  finalize(d, sums, items);
  return 0;
}

// the classification term alone (also the focal replacement pass of PPYoloELoss, csrc/focal_cls.cu), elementwise over n logits
extern "C" int cls_term_host(int focal, float alpha, const float* x, const float* q, int n, float* loss, float* grad) {
  for (int i = 0; i < n; ++i) sgb_pose::cls_term(focal, alpha, x[i], q[i], loss + i, grad + i);
  return 0;
}
