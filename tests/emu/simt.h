// TEST INFRASTRUCTURE ONLY (see cuda_runtime.h).  SIMT execution of one thread block on the host for kernels that use
// shared memory, block barriers and warp collectives: every CUDA thread of the block is a fiber (ucontext) on one OS
// thread; a fiber runs until it reaches __syncthreads / a *_sync warp primitive, deposits its operand and yields to
// the block scheduler until every live participant has arrived.  Blocks run one after another.  Semantics kept:
//   * __syncthreads releases when all threads of the block that have not exited have arrived;
//   * a warp collective with mask m completes when every lane of m that has not exited has arrived; lanes read the
//     operands of that rendezvous only (disjoint groups of one warp may rendezvous independently);
//   * nothing else synchronises: code that relies on implicit warp lockstep without __syncwarp breaks here, as it may
//     on Volta and later.
// A block in which no fiber can make progress aborts with a message (barrier divergence / deadlock).
#pragma once
#include <stdio.h>
#include <ucontext.h>

#include <functional>
#include <vector>

#if defined(__SANITIZE_ADDRESS__)
#include <sanitizer/common_interface_defs.h>
#define GSB_EMU_ASAN 1
#endif

namespace gsb_emu {

constexpr size_t FIBER_STACK = 256 * 1024;

struct Warp {
    unsigned alive = 0, arrived = 0;
    unsigned long long slot[32];
    unsigned long long res[32][32];  // res[l][m]: operand of lane m in the last rendezvous lane l took part in (private
                                     // copies: a fast group may complete its next rendezvous before a slow lane reads)
    unsigned want[32];       // mask each waiting lane called with
    unsigned gen[32] = {0};  // per-lane rendezvous generation
    unsigned res_mask[32];   // participants of that rendezvous
};

struct Block {
    int n = 0, live = 0, cur = -1;
    int bar_arrived = 0;
    unsigned bar_gen = 0;
    unsigned long progress = 0;    // bumped whenever shared state changes (deadlock detection)
    std::vector<ucontext_t> ctx;
    std::vector<char> done;
    std::vector<Warp> warps;
    ucontext_t sched;
    std::function<void()> body;
    dim3 block_dim;
    // AddressSanitizer has to be told about every stack switch (fake-stack handles per fiber, the scheduler's bounds)
    std::vector<void *> fake;
    void *sched_fake = nullptr;
    const void *sched_bottom = nullptr;
    size_t sched_size = 0;
};

static Block *g_block = nullptr;
static std::vector<char *> g_stacks;
static std::vector<char> g_dyn_smem;

static inline void *dyn_smem() { return g_dyn_smem.data(); }

static inline void yield() {
    Block *b = g_block;
    const int t = b->cur;
#ifdef GSB_EMU_ASAN
    __sanitizer_start_switch_fiber(&b->fake[t], b->sched_bottom, b->sched_size);
#endif
    swapcontext(&b->ctx[t], &b->sched);
#ifdef GSB_EMU_ASAN
    __sanitizer_finish_switch_fiber(b->fake[t], &b->sched_bottom, &b->sched_size);
#endif
}

static void fiber_main() {
    Block *b = g_block;
#ifdef GSB_EMU_ASAN
    __sanitizer_finish_switch_fiber(nullptr, &b->sched_bottom, &b->sched_size);
#endif
    b->body();
    const int t = b->cur;
    b->done[t] = 1;
    b->live--;
    b->warps[t >> 5].alive &= ~(1u << (t & 31));
    b->progress++;
#ifdef GSB_EMU_ASAN
    __sanitizer_start_switch_fiber(nullptr, b->sched_bottom, b->sched_size);   // nullptr: this fiber's fake stack dies
#endif
    swapcontext(&b->ctx[t], &b->sched);   // never resumed
}

static inline void set_thread_index(const Block &b, int t) {
    threadIdx.x = t % b.block_dim.x;
    threadIdx.y = (t / b.block_dim.x) % b.block_dim.y;
    threadIdx.z = t / (b.block_dim.x * b.block_dim.y);
}

static inline void run_block(Block &b) {
    g_block = &b;
    const int n = b.n;
    while ((int)g_stacks.size() < n) g_stacks.push_back((char *)malloc(FIBER_STACK));
    b.ctx.resize(n);
    b.fake.assign(n, nullptr);
    b.done.assign(n, 0);
    b.warps.assign((n + 31) / 32, Warp());
    b.live = n;
    b.bar_arrived = 0;
    for (int t = 0; t < n; ++t) {
        getcontext(&b.ctx[t]);
        b.ctx[t].uc_stack.ss_sp = g_stacks[t];
        b.ctx[t].uc_stack.ss_size = FIBER_STACK;
        b.ctx[t].uc_link = nullptr;
        makecontext(&b.ctx[t], fiber_main, 0);
        b.warps[t >> 5].alive |= 1u << (t & 31);
    }
    // Resume order inside a round: any order must give the same results for code whose only inter-thread ordering comes
    // from barriers and warp collectives.  GSB_EMU_ORDER=reverse | random[:seed] turns the scheduler into a race fuzzer
    // (tests/test_simt_order_cpu.py): a missing __syncwarp / __syncthreads shows up as a result that depends on it.
    static const char *order_env = getenv("GSB_EMU_ORDER");
    static unsigned long long rng = order_env && !strncmp(order_env, "random", 6)
                                        ? (strlen(order_env) > 7 ? strtoull(order_env + 7, nullptr, 10) : 1ull) * 2654435761ull + 1
                                        : 0ull;
    const bool reverse = order_env && !strcmp(order_env, "reverse");
    std::vector<int> order(n);
    for (int t = 0; t < n; ++t) order[t] = reverse ? n - 1 - t : t;
    int idle_rounds = 0;
    while (b.live > 0) {
        const unsigned long before = b.progress;
        if (rng)
            for (int i = n - 1; i > 0; --i) {   // Fisher-Yates with an xorshift generator, a new permutation every round
                rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
                std::swap(order[i], order[rng % (unsigned long long)(i + 1)]);
            }
        for (int k = 0; k < n; ++k) {
            const int t = order[k];
            if (b.done[t]) continue;
            b.cur = t;
            set_thread_index(b, t);
#ifdef GSB_EMU_ASAN
            __sanitizer_start_switch_fiber(&b.sched_fake, g_stacks[t], FIBER_STACK);
#endif
            swapcontext(&b.sched, &b.ctx[t]);
#ifdef GSB_EMU_ASAN
            __sanitizer_finish_switch_fiber(b.sched_fake, nullptr, nullptr);
#endif
        }
        if (b.progress == before) {
            if (++idle_rounds > 2) {
                fprintf(stderr, "gsb_emu: block (%u,%u) cannot make progress: %d live threads, %d at the block barrier "
                        "(divergent barrier or a warp collective naming a lane that never arrives)\n",
                        blockIdx.x, blockIdx.y, b.live, b.bar_arrived);
                abort();
            }
        } else {
            idle_rounds = 0;
        }
    }
    g_block = nullptr;
}

static inline void syncthreads() {
    Block *b = g_block;
    const unsigned gen = b->bar_gen;
    b->bar_arrived++;
    b->progress++;
    while (b->bar_gen == gen) {
        if (b->bar_arrived >= b->live) {
            b->bar_arrived = 0;
            b->bar_gen++;
            b->progress++;
            break;
        }
        yield();
    }
}

// Rendezvous of the lanes in `mask` (of the calling thread's warp); returns the warp record with
// res[lane][] holding the operands of exactly this rendezvous and *participants the lanes that took part.
static inline Warp &warp_rendezvous(unsigned mask, unsigned long long value, unsigned *participants) {
    Block *b = g_block;
    const int t = b->cur, lane = t & 31;
    Warp &w = b->warps[t >> 5];
    w.slot[lane] = value;
    w.want[lane] = mask;
    w.arrived |= 1u << lane;
    b->progress++;
    const unsigned gen = w.gen[lane];
    while (w.gen[lane] == gen) {
        const unsigned eff = mask & w.alive;
        bool ready = (w.arrived & eff) == eff;
        if (ready)
            for (int l = 0; l < 32; ++l)
                if (((eff >> l) & 1u) && w.want[l] != mask) {
                    fprintf(stderr, "gsb_emu: lanes of one warp collective disagree on the mask (%08x vs %08x)\n", mask,
                            w.want[l]);
                    abort();
                }
        if (ready) {
            for (int l = 0; l < 32; ++l) {
                if (!((eff >> l) & 1u)) continue;
                for (int m = 0; m < 32; ++m)
                    if ((eff >> m) & 1u) w.res[l][m] = w.slot[m];
                w.res_mask[l] = eff;
                w.gen[l]++;
            }
            w.arrived &= ~eff;
            b->progress++;
            break;
        }
        yield();
    }
    *participants = w.res_mask[lane];
    return w;
}

template <class T> static inline unsigned long long to_bits(T v) {
    static_assert(sizeof(T) <= 8, "warp operands up to 8 bytes");
    unsigned long long u = 0;
    memcpy(&u, &v, sizeof(T));
    return u;
}
template <class T> static inline T from_bits(unsigned long long u) {
    T v;
    memcpy(&v, &u, sizeof(T));
    return v;
}

static inline int lane_id() { return g_block->cur & 31; }

template <class T> static inline T shfl(unsigned mask, T v, int src_lane) {
    unsigned part;
    // every lane must leave the rendezvous with its partner's operand of THIS rendezvous: read before returning
    Warp &w = warp_rendezvous(mask, to_bits(v), &part);
    if (src_lane < 0 || src_lane > 31) return v;          // out of range: the lane keeps its own value (defined)
    if (!((part >> src_lane) & 1u)) {
        // reading a lane that is not part of the collective (exited, or outside the mask) is undefined in CUDA
        fprintf(stderr, "gsb_emu: lane %d of block (%u,%u) shuffles from lane %d, which is not a participant "
                "(mask %08x, participants %08x): undefined behaviour on the GPU\n", lane_id(), blockIdx.x, blockIdx.y,
                src_lane, mask, part);
        abort();
    }
    return from_bits<T>(w.res[lane_id()][src_lane]);
}

template <class F> struct SimtLauncher {
    dim3 grid, block;
    size_t smem;
    F f;
    template <class... A> void operator()(A... a) const {
        gridDim = grid;
        blockDim = block;
        if (g_dyn_smem.size() < smem + 16) g_dyn_smem.resize(smem + 16);
        Block b;
        b.n = (int)(block.x * block.y * block.z);
        b.block_dim = block;
        b.body = [&]() { f(a...); };
        for (unsigned z = 0; z < grid.z; ++z)
            for (unsigned y = 0; y < grid.y; ++y)
                for (unsigned x = 0; x < grid.x; ++x) {
                    blockIdx.x = x; blockIdx.y = y; blockIdx.z = z;
                    run_block(b);
                }
    }
};
template <class F> static inline SimtLauncher<F> launch(dim3 grid, dim3 block, size_t smem, F f) {
    return SimtLauncher<F>{grid, block, smem, f};
}

}  // namespace gsb_emu

static inline void __syncthreads() { gsb_emu::syncthreads(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) {
    unsigned part;
    gsb_emu::warp_rendezvous(mask, 0, &part);
}
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src, int = 32) { return gsb_emu::shfl(mask, v, src); }
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int o, int = 32) {
    return gsb_emu::shfl(mask, v, gsb_emu::lane_id() ^ o);
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int = 32) {
    return gsb_emu::shfl(mask, v, gsb_emu::lane_id() - (int)d);
}
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int = 32) {
    return gsb_emu::shfl(mask, v, gsb_emu::lane_id() + (int)d);
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    unsigned part;
    gsb_emu::Warp &w = gsb_emu::warp_rendezvous(mask, pred ? 1ull : 0ull, &part);
    const int me = gsb_emu::lane_id();
    unsigned r = 0;
    for (int l = 0; l < 32; ++l)
        if (((part >> l) & 1u) && w.res[me][l]) r |= 1u << l;
    return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {      // REDUX.OR (sm_80+)
    unsigned part;
    gsb_emu::Warp &w = gsb_emu::warp_rendezvous(mask, (unsigned long long)v, &part);
    const int me = gsb_emu::lane_id();
    unsigned r = 0;
    for (int l = 0; l < 32; ++l)
        if ((part >> l) & 1u) r |= (unsigned)w.res[me][l];
    return r;
}
static inline int __reduce_max_sync(unsigned mask, int v) {               // REDUX.MAX.S32
    unsigned part;
    gsb_emu::Warp &w = gsb_emu::warp_rendezvous(mask, (unsigned long long)(unsigned)v, &part);
    const int me = gsb_emu::lane_id();
    int r = v;
    for (int l = 0; l < 32; ++l)
        if ((part >> l) & 1u) { const int x = (int)(unsigned)w.res[me][l]; r = x > r ? x : r; }
    return r;
}
static inline int __all_sync(unsigned mask, int pred) {
    unsigned part;
    gsb_emu::Warp &w = gsb_emu::warp_rendezvous(mask, pred ? 1ull : 0ull, &part);
    const int me = gsb_emu::lane_id();
    for (int l = 0; l < 32; ++l)
        if (((part >> l) & 1u) && !w.res[me][l]) return 0;
    return 1;
}
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
