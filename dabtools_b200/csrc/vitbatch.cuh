// vitbatch.cuh -- host-side assembly of Viterbi work: jobs -> length-sorted warp groups.
#pragma once
#include <algorithm>
#include <vector>

#include "viterbi.cuh"

namespace dabgpu {

// Collects codewords, packs them 32 per warp by length (longest first, so the block
// scheduler starts the long ones early), uploads the descriptors and launches.
struct VitBatch {
  std::vector<VitJob> jobs;
  std::vector<VitJob> sorted;
  std::vector<VitGroup> groups;
  DevBuf d_jobs, d_groups, d_dec;
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
  // build groups; returns number of groups
  int plan() {
    sorted = jobs;
    std::stable_sort(sorted.begin(), sorted.end(),
                     [](const VitJob &a, const VitJob &b) { return a.nbits > b.nbits; });
    groups.clear();
    dec_words = 0;
    size_t i = 0;
    while (i < sorted.size()) {
      size_t j = i;
      while (j < sorted.size() && j - i < 32 && sorted[j].nbits == sorted[i].nbits) j++;
      VitGroup g;
      g.job0 = (uint32_t)i;
      g.nlanes = (uint32_t)(j - i);
      g.nsteps = sorted[i].nbits + 6;
      g.pad = 0;
      g.dec_off = dec_words;
      dec_words += vit_group_dec_words(g.nsteps);
      groups.push_back(g);
      i = j;
    }
    return (int)groups.size();
  }
  // upload descriptors (pinned staging, async) and launch on `st`
  std::vector<VitJob> planned;  // the job list the device descriptors were built from
  int run(const uint8_t *d_steps, uint8_t *d_out, cudaStream_t st) {
    if (jobs.empty()) return DABGPU_OK;
    // steady state: identical job list as last time -> descriptors on the device are still valid
    if (planned.size() == jobs.size() && d_jobs.p &&
        memcmp(planned.data(), jobs.data(), jobs.size() * sizeof(VitJob)) == 0)
      return launch_viterbi(d_steps, d_out, d_dec.as<uint2>(), d_jobs.as<VitJob>(), d_groups.as<VitGroup>(),
                            (int)groups.size(), st);
    plan();
    planned = jobs;
    int rc;
    const size_t jb = sorted.size() * sizeof(VitJob), gb = groups.size() * sizeof(VitGroup);
    if ((rc = d_jobs.reserve(jb + sizeof(VitJob) * 32))) return rc;  // lanes may over-read job0+lane
    if ((rc = d_groups.reserve(gb))) return rc;
    if ((rc = d_dec.reserve(dec_words * sizeof(uint2)))) return rc;
    // the staging buffer may still be in flight from the previous run on this stream
    CUDA_TRY(cudaStreamSynchronize(st));
    if ((rc = h_stage.reserve(jb + gb))) return rc;
    memcpy(h_stage.p, sorted.data(), jb);
    memcpy((char *)h_stage.p + jb, groups.data(), gb);
    CUDA_TRY(cudaMemcpyAsync(d_jobs.p, h_stage.p, jb, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_groups.p, (char *)h_stage.p + jb, gb, cudaMemcpyHostToDevice, st));
    return launch_viterbi(d_steps, d_out, d_dec.as<uint2>(), d_jobs.as<VitJob>(), d_groups.as<VitGroup>(),
                          (int)groups.size(), st);
  }
  void release() {
    d_jobs.release();
    d_groups.release();
    d_dec.release();
    h_stage.release();
  }
};

}  // namespace dabgpu
