// emul_runtime.cpp -- TEST INFRASTRUCTURE (see cuda_shim.h): the globals of the host re-interpretation and the block runner.
#include "cuda_shim.h"

thread_local GsbEmulIdx threadIdx, blockIdx;
GsbEmulIdx gridDim, blockDim;
GsbEmulBlock *g_emul_block = nullptr;

void gsb_emul_run_block(unsigned threads, size_t dyn_smem, const std::function<void()> &body) {
    GsbEmulBlock ctx;
    ctx.threads = threads;
    ctx.block_bar = std::make_unique<std::barrier<>>(threads);
    const unsigned warps = (threads + 31) / 32;
    for (unsigned w = 0; w < warps; ++w) ctx.warp_bar.emplace_back(std::make_unique<std::barrier<>>(std::min(32u, threads - 32 * w)));
    ctx.xchg.assign((size_t)warps * 32, 0);
    ctx.dyn_smem.assign(dyn_smem + 64, 0);
    g_emul_block = &ctx;
    std::vector<std::thread> th;
    th.reserve(threads);
    for (unsigned t = 0; t < threads; ++t)
        th.emplace_back([&ctx, &body, t] {
            threadIdx.x = t; threadIdx.y = 0; threadIdx.z = 0;
            body();
            // a thread that has returned counts as arrived at every later barrier, as on the device
            ctx.xchg[t] = 0;
            ctx.warp_bar[t / 32]->arrive_and_drop();
            ctx.block_bar->arrive_and_drop();
        });
    for (auto &x : th) x.join();
    g_emul_block = nullptr;
}
