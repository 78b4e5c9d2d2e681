// simt_emu.cc — scheduler of the test-only SIMT emulator (see simt_emu.h).
#include <execinfo.h>
#include <signal.h>
#include <stdarg.h>
#include <unistd.h>

#include "simt_emu.h"

namespace emu {

State S;
uint64_t stats[16];
static void print_stats() {
    if (!getenv("KC_EMU_STATS")) return;
    fprintf(stderr, "simt_emu stats:");
    for (int i = 0; i < 16; i++) fprintf(stderr, " [%d]=%llu", i, (unsigned long long)stats[i]);
    fprintf(stderr, " launches=%llu switches=%llu\n", (unsigned long long)S.launches, (unsigned long long)S.switches);
}
dim3 threadIdx_, blockIdx_, blockDim_, gridDim_;

static constexpr size_t STACK_BYTES = 256 * 1024;

[[noreturn]] void fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    fprintf(stderr, "simt_emu: ");
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    if (S.in_kernel)
        fprintf(stderr, "  [block (%u,%u) thread %u]", blockIdx_.x, blockIdx_.y, threadIdx_.x);
    fprintf(stderr, "\n");
    abort();
}

extern "C" void emu_ctx_switch(void** from_sp, void* to_sp);
asm(R"(
.text
.globl emu_ctx_switch
.type emu_ctx_switch,@function
emu_ctx_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_ctx_switch,.-emu_ctx_switch
)");

static inline uint64_t next_rand() {
    uint64_t z = (S.rng += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static void set_ids(int f) {
    threadIdx_.x = (unsigned)f % S.block.x;
    threadIdx_.y = ((unsigned)f / S.block.x) % S.block.y;
    threadIdx_.z = (unsigned)f / (S.block.x * S.block.y);
}

static void to_scheduler() {
    Fiber& me = S.fibers[S.cur];
    S.switches++;
    emu_ctx_switch(&me.sp, S.sched_sp);
}

void make_runnable(int f) {
    Fiber& fb = S.fibers[f];
    if (fb.state != 1) fail("internal: waking a fiber that is not blocked");
    fb.state = 0;
    fb.run_pos = (int)S.runnable.size();
    S.runnable.push_back(f);
}

static void remove_runnable(int f) {
    Fiber& fb = S.fibers[f];
    const int pos = fb.run_pos;
    const int last = S.runnable.back();
    S.runnable[pos] = last;
    S.fibers[last].run_pos = pos;
    S.runnable.pop_back();
    fb.run_pos = -1;
}

void yield_blocked() {
    Fiber& me = S.fibers[S.cur];
    me.state = 1;
    remove_runnable(S.cur);
    to_scheduler();
}

void maybe_preempt() {
    if (!S.in_kernel || S.preempt_mask == 0) return;
    if ((next_rand() & S.preempt_mask) != 0) return;
    to_scheduler();  // stays runnable
}

void maybe_preempt_always() {
    if (S.in_kernel) to_scheduler();
}

static void check_barrier_release() {
    if (S.bar_arrived > 0 && S.bar_arrived == S.live) {
        S.bar_arrived = 0;
        std::vector<int> w;
        w.swap(S.bar_waiters);
        for (int f : w) make_runnable(f);
    }
}

void syncthreads() {
    const int me = S.cur;
    S.bar_arrived++;
    if (S.bar_arrived == S.live) {  // last one in: release the others, keep running
        S.bar_arrived = 0;
        std::vector<int> w;
        w.swap(S.bar_waiters);
        for (int f : w) make_runnable(f);
        return;
    }
    S.bar_waiters.push_back(me);
    yield_blocked();
}

uint32_t warp_exchange(uint32_t mask, uint32_t v, int kind, int arg) {
    const int f = S.cur;
    const int tpb = (int)(S.block.x * S.block.y * S.block.z);
    const int lane = f & 31, wid = f >> 5;
    Warp& W = S.warps[wid];
    // lanes that exist in this warp (the last warp of a block may be partial)
    const int lanes_here = std::min(32, tpb - wid * 32);
    const uint32_t exist = lanes_here == 32 ? 0xffffffffu : ((1u << lanes_here) - 1u);
    mask &= exist;
    if (!(mask & (1u << lane))) fail("warp collective called by lane %d which is not in mask 0x%08x", lane, mask);
    const uint32_t p = W.gen & 1u;
    W.vals[p][lane] = v;
    if (kind == EMU_BALLOT) {
        if (W.arrived == 0) W.ballot[p] = 0;
        if (v) W.ballot[p] |= 1u << lane;
    }
    W.arrived |= 1u << lane;
    if ((W.arrived & mask) == mask) {
        if (W.arrived != mask) fail("warp collective: lanes outside the mask arrived (mask 0x%08x arrived 0x%08x)", mask, W.arrived);
        W.arrived = 0;
        W.gen++;
        for (int l = 0; l < 32; l++)
            if ((mask & (1u << l)) && l != lane) make_runnable(wid * 32 + l);
    } else {
        yield_blocked();
    }
    switch (kind) {
        case EMU_SHFL_IDX: {
            const int src = arg & 31;
            return (mask & (1u << src)) ? W.vals[p][src] : v;
        }
        case EMU_SHFL_DOWN: {
            const int src = lane + arg;
            return (src < 32 && (mask & (1u << src))) ? W.vals[p][src] : v;
        }
        case EMU_SHFL_UP: {
            const int src = lane - arg;
            return (src >= 0 && (mask & (1u << src))) ? W.vals[p][src] : v;
        }
        case EMU_SHFL_XOR: {
            const int src = lane ^ arg;
            return (src < 32 && (mask & (1u << src))) ? W.vals[p][src] : v;
        }
        case EMU_BALLOT:
            return W.ballot[p] & mask;
        case 7: {  // EMU_MATCH_ANY
            uint32_t m = 0;
            for (int l = 0; l < 32; l++)
                if ((mask & (1u << l)) && W.vals[p][l] == v) m |= 1u << l;
            return m;
        }
        case EMU_REDUCE_ADD: {
            uint32_t sum = 0;
            for (int l = 0; l < 32; l++)
                if (mask & (1u << l)) sum += W.vals[p][l];
            return sum;
        }
        default:
            return 0;
    }
}

static void fiber_main() {
    (*S.body)();
    Fiber& me = S.fibers[S.cur];
    me.state = 2;
    remove_runnable(S.cur);
    S.live--;
    check_barrier_release();
    to_scheduler();
    fail("internal: finished fiber resumed");
}

static void prepare_fiber(Fiber& fb) {
    if (!fb.stack) {
        fb.stack = (char*)mmap(nullptr, STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (fb.stack == MAP_FAILED) fail("cannot map a fiber stack");
        mprotect(fb.stack, 4096, PROT_NONE);  // overflow guard
    }
    uintptr_t top = ((uintptr_t)fb.stack + STACK_BYTES) & ~(uintptr_t)15;
    void** sp = (void**)top;
    *--sp = nullptr;               // fake return address of fiber_main (keeps the ABI alignment)
    *--sp = (void*)&fiber_main;    // 'ret' target of the first switch
    for (int i = 0; i < 6; i++) *--sp = nullptr;  // rbp rbx r12 r13 r14 r15
    fb.sp = sp;
    fb.state = 0;
}

void describe_address(const void* addr);

static void on_segv(int, siginfo_t* si, void*) {
    static char buf[512];
    int n = snprintf(buf, sizeof buf, "simt_emu: SIGSEGV at address %p  [block (%u,%u) thread %u]\n", si->si_addr, blockIdx_.x,
                     blockIdx_.y, threadIdx_.x);
    if (write(2, buf, n) < 0) {}
    describe_address(si->si_addr);
    void* bt[32];
    const int m = backtrace(bt, 32);
    backtrace_symbols_fd(bt, m, 2);
    _exit(139);
}

static void install_segv_handler() {
    static char altstack[1 << 16];
    stack_t ss;
    ss.ss_sp = altstack;
    ss.ss_size = sizeof altstack;
    ss.ss_flags = 0;
    sigaltstack(&ss, nullptr);
    struct sigaction sa;
    memset(&sa, 0, sizeof sa);
    sa.sa_sigaction = on_segv;
    sa.sa_flags = SA_SIGINFO | SA_ONSTACK;
    sigaction(SIGSEGV, &sa, nullptr);
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    if (S.in_kernel) fail("nested launch");
    const int tpb = (int)(block.x * block.y * block.z);
    if (tpb < 1 || tpb > 1024) fail("bad block size %d", tpb);
    if (smem > 227 * 1024) fail("dynamic shared memory %zu exceeds 227 KB", smem);
    static bool init = false;
    static uint64_t seed = 0;
    if (!init) {
        init = true;
        install_segv_handler();
        atexit(print_stats);
        const char* s = getenv("KC_EMU_SEED");
        seed = s ? strtoull(s, nullptr, 0) : 0;
        const char* p = getenv("KC_EMU_PREEMPT_SHIFT");  // yield with probability 2^-shift at every preemption point
        const int shift = p ? atoi(p) : 3;
        S.preempt_mask = seed ? ((1u << shift) - 1u) : 0u;
        if (seed && shift == 0) S.preempt_mask = 0;  // shift 0 would mean "always": use mask 0 == never; keep it sane
    }
    S.rng = seed * 0x2545F4914F6CDD1Dull + S.launches;
    S.launches++;
    S.grid = grid;
    S.block = block;
    S.smem_bytes = smem;
    // dynamic shared memory ENDS at a guard page (16-byte granular, like the device allocations below): a kernel that
    // reads or writes past the size its host code asked for dies here with SIGSEGV instead of passing by luck
    const size_t smem_page = 4096, smem_body = (smem + 15) & ~(size_t)15;
    const size_t smem_map = ((smem_body + smem_page - 1) & ~(smem_page - 1)) + smem_page;
    char* smem_base = (char*)mmap(nullptr, smem_map, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (smem_base == (char*)MAP_FAILED) fail("mmap of %zu bytes of shared memory failed", smem_map);
    mprotect(smem_base + smem_map - smem_page, smem_page, PROT_NONE);
    S.dyn_smem = smem_base + smem_map - smem_page - smem_body;
    if ((int)S.fibers.size() < tpb) S.fibers.resize(tpb);
    S.warps.assign((tpb + 31) / 32, Warp());
    S.body = &body;
    gridDim_ = grid;
    blockDim_ = block;
    S.in_kernel = true;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                blockIdx_.x = bx;
                blockIdx_.y = by;
                blockIdx_.z = bz;
                memset(S.dyn_smem, 0xCD, smem);
                S.runnable.clear();
                S.bar_waiters.clear();
                S.bar_arrived = 0;
                for (auto& w : S.warps) w = Warp();
                for (int f = 0; f < tpb; f++) {
                    prepare_fiber(S.fibers[f]);
                    S.fibers[f].run_pos = f;
                    S.runnable.push_back(f);
                }
                S.live = tpb;
                size_t rr = 0;
                while (S.live > 0) {
                    if (S.runnable.empty())
                        fail("deadlock: %d threads alive, none runnable (divergent barrier or collective?)", S.live);
                    int f;
                    if (seed) {
                        f = S.runnable[next_rand() % S.runnable.size()];
                    } else {
                        if (rr >= S.runnable.size()) rr = 0;
                        f = S.runnable[rr];
                    }
                    S.cur = f;
                    set_ids(f);
                    emu_ctx_switch(&S.sched_sp, S.fibers[f].sp);
                    // back in the scheduler: without a seed keep running the same warp's
                    // neighbours in order (cheap collectives); position rr now holds another fiber
                    if (!seed && S.fibers[f].state == 0) rr++;
                }
            }
    S.in_kernel = false;
    S.cur = -1;
    munmap(smem_base, smem_map);
    S.dyn_smem = nullptr;
    S.smem_bytes = 0;
}

}  // namespace emu

// ---- device memory with a guard page behind every allocation ---------------------------
namespace {
struct Alloc {
    void* map;
    size_t map_bytes;
};
std::map<void*, Alloc>& allocs() {
    static std::map<void*, Alloc> a;
    return a;
}
}  // namespace

namespace emu {
void describe_address(const void* addr) {
    char buf[256];
    for (auto& kv : allocs()) {
        const char* m = (const char*)kv.second.map;
        if ((const char*)addr >= m && (const char*)addr < m + kv.second.map_bytes) {
            const long off = (const char*)addr - (const char*)kv.first;
            const int n = snprintf(buf, sizeof buf, "  inside the mapping of device allocation %p: offset %ld from its start (guard page behind its end)\n",
                                   kv.first, off);
            if (write(2, buf, n) < 0) {}
            return;
        }
    }
    const int n = snprintf(buf, sizeof buf, "  not near any device allocation\n");
    if (write(2, buf, n) < 0) {}
}
}  // namespace emu

cudaError_t emu_cuda_malloc(void** p, size_t n) {
    const size_t page = 4096;
    const size_t body = (n + 15) & ~(size_t)15;
    const size_t map_bytes = ((body + page - 1) & ~(page - 1)) + page;
    void* m = mmap(nullptr, map_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (m == MAP_FAILED) {
        *p = nullptr;
        return cudaErrorMemoryAllocation;
    }
    char* guard = (char*)m + map_bytes - page;
    mprotect(guard, page, PROT_NONE);
    char* user = guard - body;  // the allocation ends (16-byte granular) at the guard page
    memset(user, 0xA5, body);   // device memory is not zero-initialised
    allocs()[user] = Alloc{m, map_bytes};
    *p = user;
    return cudaSuccess;
}

cudaError_t cudaFree(void* p) {
    if (!p) return cudaSuccess;
    auto it = allocs().find(p);
    if (it == allocs().end()) emu::fail("cudaFree of an unknown pointer %p", p);
    munmap(it->second.map, it->second.map_bytes);
    allocs().erase(it);
    return cudaSuccess;
}

cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    memset(p, 0, sizeof *p);
    p->major = 10;
    p->minor = 0;
    const char* s = getenv("KC_EMU_SMS");
    p->multiProcessorCount = s ? atoi(s) : 4;
    p->sharedMemPerBlockOptin = 227 * 1024;
    snprintf(p->name, sizeof p->name, "simt_emu (CPU, test only)");
    return cudaSuccess;
}
