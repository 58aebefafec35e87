// Runtime of the CUDA-semantics emulator (see cuda_emu.h).  TEST INFRASTRUCTURE ONLY.
#include "cuda_emu.h"

namespace emu {

thread_local ThreadCtx ctx;

static std::vector<unsigned char> g_smem;
static std::unique_ptr<std::barrier<>> g_block_bar;
static std::vector<std::unique_ptr<std::barrier<>>> g_warp_bar;
static float g_shfl[64][32];

unsigned char* dyn_smem() { return g_smem.data(); }

void sync_block() {
    if (!ctx.threaded) {
        fprintf(stderr, "emu: __syncthreads() inside a kernel launched with FDN_LAUNCH_SEQ\n");
        abort();
    }
    g_block_bar->arrive_and_wait();
}

float shfl(float v, int src_lane) {
    if (!ctx.threaded) {
        fprintf(stderr, "emu: warp shuffle inside a kernel launched with FDN_LAUNCH_SEQ\n");
        abort();
    }
    int w = ctx.linear_tid >> 5, l = ctx.linear_tid & 31;
    g_shfl[w][l] = v;
    g_warp_bar[w]->arrive_and_wait();
    float r = g_shfl[w][src_lane & 31];
    g_warp_bar[w]->arrive_and_wait();
    return r;
}

namespace {
struct Pool {
    std::vector<std::thread> threads;
    std::mutex m;
    std::condition_variable cv, cv_done;
    uint64_t gen = 0;
    int done = 0;
    // job
    const std::function<void()>* body = nullptr;
    dim3 grid, block;
    uint3 bid;
    int nthreads = 0;

    void worker(int id, uint64_t seen) {
        for (;;) {
            std::unique_lock<std::mutex> lk(m);
            cv.wait(lk, [&] { return gen != seen; });
            seen = gen;
            bool active = id < nthreads;
            const std::function<void()>* b = body;
            dim3 g = grid, bd = block;
            uint3 bi = bid;
            lk.unlock();
            if (active) {
                ctx.threaded = true;
                ctx.linear_tid = id;
                ctx.tid.x = id % bd.x;
                ctx.tid.y = (id / bd.x) % bd.y;
                ctx.tid.z = id / (bd.x * bd.y);
                ctx.bid = bi;
                ctx.bdim = bd;
                ctx.gdim = g;
                (*b)();
                g_warp_bar[id >> 5]->arrive_and_drop();
                g_block_bar->arrive_and_drop();
            }
            lk.lock();
            if (++done == (int)threads.size()) cv_done.notify_one();
        }
    }
    void ensure(int n) {
        std::unique_lock<std::mutex> lk(m);
        while ((int)threads.size() < n) {
            int id = (int)threads.size();
            uint64_t g0 = gen;
            threads.emplace_back([this, id, g0] { worker(id, g0); });
            threads.back().detach();
        }
    }
    void run_block(const std::function<void()>& b, dim3 g, dim3 bd, uint3 bi, int nt) {
        std::unique_lock<std::mutex> lk(m);
        body = &b;
        grid = g;
        block = bd;
        bid = bi;
        nthreads = nt;
        done = 0;
        ++gen;
        cv.notify_all();
        cv_done.wait(lk, [&] { return done == (int)threads.size(); });
    }
};
Pool& g_pool = *new Pool;   // leaked on purpose: detached workers may still wait on its cv at process exit
}  // namespace

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body, bool threaded) {
    int nt = (int)(block.x * block.y * block.z);
    if (nt <= 0 || nt > 1024) {
        fprintf(stderr, "emu: bad block size %d\n", nt);
        abort();
    }
    if (g_smem.size() < smem_bytes + 64) g_smem.resize(smem_bytes + 64);
    if (threaded) g_pool.ensure(nt);
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                uint3 bi{bx, by, bz};
                if (threaded) {
                    g_block_bar.reset(new std::barrier<>(nt));
                    g_warp_bar.clear();
                    for (int w = 0; w * 32 < nt; ++w)
                        g_warp_bar.emplace_back(new std::barrier<>(std::min(32, nt - w * 32)));
                    g_pool.run_block(body, grid, block, bi, nt);
                } else {
                    ThreadCtx saved = ctx;
                    ctx.threaded = false;
                    ctx.bid = bi;
                    ctx.bdim = block;
                    ctx.gdim = grid;
                    for (int id = 0; id < nt; ++id) {
                        ctx.linear_tid = id;
                        ctx.tid.x = id % block.x;
                        ctx.tid.y = (id / block.x) % block.y;
                        ctx.tid.z = id / (block.x * block.y);
                        body();
                    }
                    ctx = saved;
                }
            }
}

}  // namespace emu
