// vitbatch.cuh -- host-side assembly of Viterbi work: jobs -> equal-length warp groups ->
// per-warp-scheduler work lists for the persistent decoder kernel.
#pragma once
#include <algorithm>
#include <vector>

#include "viterbi.cuh"

namespace dabgpu {

// Collects codewords, packs them 32 per warp by length (longest first: the persistent kernel's
// warps pull groups in that order, which is list scheduling with the shortest jobs last), uploads
// the descriptors and launches.  The plan is cached: an identical job list
// (the steady state of a locked receiver) re-launches without any host work or upload.
struct VitBatch {
  std::vector<VitJob> jobs;
  std::vector<VitJob> sorted;
  std::vector<VitJob> planned;  // the job list the device descriptors were built from
  std::vector<VitGroup> groups;
  DevBuf d_jobs, d_groups, d_queue, d_dec;
  PinBuf h_stage;
  uint64_t dec_words = 0;
  uint64_t total_steps = 0;  // sum over codewords of nbits+6 (for ACS/s accounting)

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
    groups.swap(g0);  // longest first: the order in which the kernel's warps pull them
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
      const size_t jb = sorted.size() * sizeof(VitJob), gb = groups.size() * sizeof(VitGroup);
      if ((rc = d_jobs.reserve(jb))) return rc;
      if ((rc = d_groups.reserve(gb))) return rc;
      if ((rc = d_queue.reserve(64))) return rc;
      if ((rc = d_dec.reserve(dec_words * sizeof(uint2)))) return rc;
      // the staging buffer may still be in flight from the previous upload on this stream
      CUDA_TRY(cudaStreamSynchronize(st));
      if ((rc = h_stage.reserve(jb + gb))) return rc;
      memcpy(h_stage.p, sorted.data(), jb);
      memcpy((char *)h_stage.p + jb, groups.data(), gb);
      CUDA_TRY(cudaMemcpyAsync(d_jobs.p, h_stage.p, jb, cudaMemcpyHostToDevice, st));
      CUDA_TRY(cudaMemcpyAsync(d_groups.p, (char *)h_stage.p + jb, gb, cudaMemcpyHostToDevice, st));
    }
    return relaunch(d_steps, d_out, st);
  }
  // launch again with the descriptors of the last run() (identical job list by construction)
  int relaunch(const uint8_t *d_steps, uint8_t *d_out, cudaStream_t st) {
    return launch_viterbi(d_steps, d_out, d_dec.as<uint2>(), d_jobs.as<VitJob>(), d_groups.as<VitGroup>(),
                          (int)groups.size(), d_queue.as<uint32_t>(), st);
  }
  void release() {
    d_jobs.release();
    d_groups.release();
    d_queue.release();
    d_dec.release();
    h_stage.release();
    planned.clear();
  }
};

}  // namespace dabgpu
