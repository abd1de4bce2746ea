// vitbatch.cuh -- host-side assembly of Viterbi work: jobs -> equal-length warp groups ->
// per-warp-scheduler work lists for the persistent decoder kernel.
#pragma once
#include <algorithm>
#include <queue>
#include <vector>

#include "viterbi.cuh"

namespace dabgpu {

// Collects codewords, packs them 32 per warp by length, distributes the groups over the GPU's
// warp schedulers with longest-processing-time-first (so every scheduler sees the same number of
// trellis steps), uploads the descriptors and launches.  The plan is cached: an identical job list
// (the steady state of a locked receiver) re-launches without any host work or upload.
struct VitBatch {
  std::vector<VitJob> jobs;
  std::vector<VitJob> sorted;
  std::vector<VitJob> planned;  // the job list the device descriptors were built from
  std::vector<VitGroup> groups;
  std::vector<uint32_t> bin_start;
  DevBuf d_jobs, d_groups, d_bins, d_dec;
  PinBuf h_stage;
  uint64_t dec_words = 0;
  uint64_t total_steps = 0;  // sum over codewords of nbits+6 (for ACS/s accounting)
  int n_ctas = 0;
  bool small_ctas = false;  // one single-warp CTA per group instead of the persistent layout
  bool one_warp_ctas = false;  // what plan() chose
  bool soft = false;           // rows are 4 symbols per step, decoded by viterbi_soft_kernel
  double reserve_scale = 1.0;  // allocate this much more than the current job list needs (the
                               // caller's ratio of a full batch to this one), so stores never grow

  void clear() {
    jobs.clear();
    total_steps = 0;
  }
  void add(uint64_t in_off, uint64_t out_off, uint32_t nbits, uint32_t flags) {
    jobs.push_back(VitJob{in_off, out_off, nbits, flags});
    total_steps += nbits + 6;
  }

  void plan() {
    sorted = jobs;
    std::stable_sort(sorted.begin(), sorted.end(),
                     [](const VitJob &a, const VitJob &b) { return a.nbits > b.nbits; });
    // 32 equal-length codewords per group, longest first
    std::vector<VitGroup> g0;
    dec_words = 0;
    for (size_t i = 0; i < sorted.size();) {
      size_t j = i;
      while (j < sorted.size() && j - i < 32 && sorted[j].nbits == sorted[i].nbits) j++;
      VitGroup g;
      g.job0 = (uint32_t)i;
      g.nlanes = (uint32_t)(j - i);
      g.nsteps = sorted[i].nbits + 6;
      g.pad = 0;
      g.dec_off = dec_words;
      dec_words += vit_group_dec_words(g.nsteps);
      g0.push_back(g);
      i = j;
    }
    // sparse batches (fewer than two groups per warp scheduler) are latency-bound: one single-warp
    // CTA per group lets the hardware spread them, whatever else is running
    one_warp_ctas = soft || small_ctas || g0.size() < (size_t)device_sm_count() * 8;
    if (one_warp_ctas) {
      groups.swap(g0);
      n_ctas = (int)groups.size();
      bin_start.resize(groups.size() + 1);
      for (size_t i = 0; i <= groups.size(); i++) bin_start[i] = (uint32_t)i;
      return;
    }
    // LPT over the CTAs' scheduler quarters (warp w of a CTA runs on scheduler w % 4); a quarter's
    // groups are then dealt LPT to its VIT_WARPS/4 warps so that they overlap each other's latencies.
    // With several CTAs per SM every CTA is planned on its own: which CTAs end up sharing an SM is the
    // hardware's choice, and equal work lists make it irrelevant.
    const int n_sm = device_sm_count() * VIT_CTAS_PER_SM;
    const int per_sched = VIT_WARPS / 4;
    const int n_sched = std::min<int>(n_sm * 4, std::max<size_t>(1, g0.size()));
    n_ctas = std::min(n_sm, n_sched);  // small batches spread over SMs before doubling up
    typedef std::pair<uint64_t, int> Load;  // (steps so far, scheduler)
    std::priority_queue<Load, std::vector<Load>, std::greater<Load>> pq;
    for (int s = 0; s < n_sched; s++) pq.push(Load(0, s));
    std::vector<std::vector<uint32_t>> sched(n_sched);
    for (size_t i = 0; i < g0.size(); i++) {  // g0 is already longest-first
      Load l = pq.top();
      pq.pop();
      sched[l.second].push_back((uint32_t)i);
      pq.push(Load(l.first + g0[i].nsteps, l.second));
    }
    const int n_bins = n_ctas * VIT_WARPS;
    std::vector<std::vector<uint32_t>> bins(n_bins);
    for (int s = 0; s < n_sched; s++) {
      // scheduler s = CTA (s % n_ctas), partition (s / n_ctas); its warps are w = part + 4 k.
      // LPT again over those warps (the list is already longest-first)
      const int cta = s % n_ctas, part = s / n_ctas;
      uint64_t load[VIT_WARPS / 4] = {};
      for (uint32_t gi : sched[s]) {
        int k = 0;
        for (int q = 1; q < per_sched; q++)
          if (load[q] < load[k]) k = q;
        load[k] += g0[gi].nsteps;
        bins[cta * VIT_WARPS + part + 4 * k].push_back(gi);
      }
    }
    groups.clear();
    bin_start.assign(n_bins + 1, 0);
    for (int b = 0; b < n_bins; b++) {
      bin_start[b] = (uint32_t)groups.size();
      for (uint32_t gi : bins[b]) groups.push_back(g0[gi]);
    }
    bin_start[n_bins] = (uint32_t)groups.size();
  }

  // upload descriptors (pinned staging, async) and launch on `st`
  int run(const uint8_t *d_steps, uint8_t *d_out, cudaStream_t st) {
    if (jobs.empty()) return DABGPU_OK;
    const bool same = planned.size() == jobs.size() && d_jobs.p &&
                      memcmp(planned.data(), jobs.data(), jobs.size() * sizeof(VitJob)) == 0;
    if (!same) {
      plan();
      planned = jobs;
      int rc;
      const size_t jb = sorted.size() * sizeof(VitJob), gb = groups.size() * sizeof(VitGroup),
                   bb = bin_start.size() * sizeof(uint32_t);
      auto full = [this](size_t bytes) { return (size_t)((double)bytes * reserve_scale) + 4096; };
      // bins are per SM; groups can be up to 32x more numerous for another job mix of the same size
      if (d_jobs.cap < jb && (rc = d_jobs.reserve(full(jb)))) return rc;
      if (d_groups.cap < gb && (rc = d_groups.reserve(full(gb) * 2))) return rc;
      if (d_bins.cap < bb && (rc = d_bins.reserve(full(bb)))) return rc;
      if (d_dec.cap < dec_words * sizeof(uint2) && (rc = d_dec.reserve(full(dec_words * sizeof(uint2)))))
        return rc;
      // the staging buffer may still be in flight from the previous upload on this stream
      CUDA_TRY(cudaStreamSynchronize(st));
      if (h_stage.cap < jb + gb + bb && (rc = h_stage.reserve(full(jb + gb + bb) + full(gb)))) return rc;
      memcpy(h_stage.p, sorted.data(), jb);
      memcpy((char *)h_stage.p + jb, groups.data(), gb);
      memcpy((char *)h_stage.p + jb + gb, bin_start.data(), bb);
      CUDA_TRY(cudaMemcpyAsync(d_jobs.p, h_stage.p, jb, cudaMemcpyHostToDevice, st));
      CUDA_TRY(cudaMemcpyAsync(d_groups.p, (char *)h_stage.p + jb, gb, cudaMemcpyHostToDevice, st));
      CUDA_TRY(cudaMemcpyAsync(d_bins.p, (char *)h_stage.p + jb + gb, bb, cudaMemcpyHostToDevice, st));
    }
    return relaunch(d_steps, d_out, st);
  }
  // launch again with the descriptors of the last run() (identical job list by construction)
  int relaunch(const uint8_t *d_steps, uint8_t *d_out, cudaStream_t st) {
    if (soft)
      return launch_viterbi_soft(d_steps, d_out, d_dec.as<uint2>(), d_jobs.as<VitJob>(), d_groups.as<VitGroup>(),
                                 (int)groups.size(), st);
    return launch_viterbi(d_steps, d_out, d_dec.as<uint2>(), d_jobs.as<VitJob>(), d_groups.as<VitGroup>(),
                          d_bins.as<uint32_t>(), n_ctas, one_warp_ctas ? 1 : VIT_WARPS, st);
  }
  void release() {
    d_jobs.release();
    d_groups.release();
    d_bins.release();
    d_dec.release();
    h_stage.release();
    planned.clear();
  }
};

}  // namespace dabgpu
