// Residency of ONE plan per device in a module's __constant__ bank (host side).
//
// Kernels on the uniform datapath read their broadcast operands from a __constant__ array
// (`LDCU`), which is a per-module global: only one plan's data can be there at a time.  A
// ConstBank instance tracks, per device, which plan is resident and which streams have kernels
// of it in flight, so that a plan switch (i) re-uploads with cudaMemcpyToSymbolAsync on the
// launching stream and (ii) orders that upload after every kernel of the previous plan, and a
// launch of the resident plan on another stream is ordered after the upload.
#pragma once

#include <mutex>

#include "pcb_common.cuh"

namespace pcb {

class ConstBank {
  public:
    static constexpr int MAX_DEV = 64;
    static constexpr int MAX_TRACKED = 16;

    // Make plan `plan_id` resident on device `d` and order `st` after its upload.  `upload(st)`
    // must enqueue the cudaMemcpyToSymbolAsync calls and return a cudaError_t.  On success the
    // device entry stays LOCKED until release() -- the caller launches on `st` in between.
    template <typename Upload>
    int acquire(int d, uint64_t plan_id, cudaStream_t st, Upload upload) {
        if (d < 0 || d >= MAX_DEV) return fail(PCB_EINVAL, "device %d out of range", d);
        Dev &R = dev_[d];
        R.m.lock();
        if (!R.uploaded && cudaEventCreateWithFlags(&R.uploaded, cudaEventDisableTiming) != cudaSuccess) {
            R.m.unlock();
            return fail(PCB_ECUDA, "cudaEventCreate failed");
        }
        if (R.plan_id != plan_id) {
            // kernels of the previous plan may still be reading the bank on other streams
            for (int i = 0; i < R.n_users; ++i)
                if (R.users[i] != st) cudaStreamWaitEvent(st, R.used[i], 0);
            const cudaError_t e = upload(st);
            if (e != cudaSuccess) {
                R.plan_id = 0;
                R.m.unlock();
                return fail(PCB_ECUDA, "upload to the constant bank failed: %s", cudaGetErrorString(e));
            }
            cudaEventRecord(R.uploaded, st);
            R.upload_stream = st;
            R.plan_id = plan_id;
            R.n_users = 0;
        } else if (st != R.upload_stream) {
            cudaStreamWaitEvent(st, R.uploaded, 0);
        }
        return PCB_OK;
    }

    // After the launch: remember that `st` has a kernel of the resident plan in flight; unlock.
    void release(int d, cudaStream_t st) {
        Dev &R = dev_[d];
        int slot = -1;
        for (int i = 0; i < R.n_users; ++i)
            if (R.users[i] == st) slot = i;
        if (slot < 0) {
            if (R.n_users == MAX_TRACKED) {
                cudaDeviceSynchronize();  // too many streams to track: device-wide fence instead
                R.n_users = 0;
            }
            slot = R.n_users++;
            R.users[slot] = st;
        }
        if (!R.used[slot]) cudaEventCreateWithFlags(&R.used[slot], cudaEventDisableTiming);
        cudaEventRecord(R.used[slot], st);
        R.m.unlock();
    }

    // A destroyed plan must not be mistaken for a later one with a recycled id.
    void forget(int d, uint64_t plan_id) {
        if (d < 0 || d >= MAX_DEV) return;
        std::lock_guard<std::mutex> lock(dev_[d].m);
        if (dev_[d].plan_id == plan_id) dev_[d].plan_id = 0;
    }

  private:
    struct Dev {
        std::mutex m;
        uint64_t plan_id = 0;
        cudaEvent_t uploaded = nullptr;
        cudaStream_t upload_stream = nullptr;
        cudaStream_t users[MAX_TRACKED];
        cudaEvent_t used[MAX_TRACKED] = {};
        int n_users = 0;
    };
    Dev dev_[MAX_DEV];
};

uint64_t next_plan_id();  // process-wide unique plan ids (pcb_abi.cu)

}  // namespace pcb
