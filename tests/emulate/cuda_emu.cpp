// cuda_emu.cpp -- fiber scheduler and runtime shim behind cuda_emu.h (TEST INFRASTRUCTURE ONLY).
#include "cuda_emu.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
#if !defined(__x86_64__)
#include <ucontext.h>
#endif

#include <deque>
#include <map>
#include <string>
#include <vector>

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

namespace emu {
namespace {

constexpr size_t STACK_BYTES = 256 * 1024;
constexpr uint64_t CANARY = 0xC0DEC0DEDEADBEEFull;

// Context switch. x86-64: a hand-written switch of the callee-saved registers and the stack pointer (ucontext's swapcontext
// makes a signal-mask system call per switch, and a sync-heavy kernel switches millions of times); elsewhere ucontext.
#if defined(__x86_64__)
struct Context { void *sp = nullptr; };
extern "C" void emu_switch(void **save_sp, void *load_sp);
asm(R"(
    .text
    .globl emu_switch
    .type emu_switch, @function
emu_switch:
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
    .size emu_switch, .-emu_switch
)");
inline void ctx_switch(Context &from, Context &to) { emu_switch(&from.sp, to.sp); }
inline void ctx_make(Context &c, unsigned char *stack, size_t bytes, void (*entry)())
{
    void **top = reinterpret_cast<void **>(stack + bytes);           // 16-byte aligned
    top[-1] = nullptr;                                               // where a return address would be: entry never returns
    top[-2] = reinterpret_cast<void *>(entry);                       // `ret` of the first switch jumps here
    for (int i = 3; i <= 8; i++) top[-i] = nullptr;                  // rbp rbx r12 r13 r14 r15
    c.sp = top - 8;
}
#else
struct Context { ucontext_t uc; };
inline void ctx_switch(Context &from, Context &to) { swapcontext(&from.uc, &to.uc); }
inline void ctx_make(Context &c, unsigned char *stack, size_t bytes, void (*entry)())
{
    getcontext(&c.uc);
    c.uc.uc_stack.ss_sp = stack;
    c.uc.uc_stack.ss_size = bytes;
    c.uc.uc_link = nullptr;
    makecontext(&c.uc, entry, 0);
}
#endif

struct Fiber {
    Context ctx;
    uint3 tid;
    unsigned linear = 0;
    bool done = false;
};

struct Warp {
    int live = 0, arrived = 0, created = 0;
    unsigned gen = 0;
    unsigned live_mask = 0;
    uint64_t pad[32] = {0};
    uint64_t res[2][32] = {{0}};
    std::vector<int> waiters;
};

struct Block {
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    std::deque<int> ready;
    int live = 0, bar_arrived = 0, bar_pred = 0;
    unsigned bar_gen = 0;
    int bar_res[2] = {0, 0};
    std::vector<int> bar_waiters;
    unsigned char *smem = nullptr;
    size_t smem_bytes = 0;
};

Context g_sched;
Block *g_blk = nullptr;
int g_cur = -1;
const std::function<void()> *g_body = nullptr;
unsigned char *g_stacks = nullptr;
size_t g_stack_count = 0;
std::string g_error;
cudaError_t g_last = cudaSuccess;
Stats g_stats = {};
// "CUDA graphs": while a capture is open, launches / memsets / copies are recorded (and executed: harmless, the real capture
// only records); a replay runs the recorded nodes again with the argument values they were recorded with.
struct GraphNode { int kind; dim3 grid, block; size_t smem; std::function<void()> body; void *dst; const void *src; int value; size_t bytes; };
std::vector<std::vector<GraphNode>> g_graphs;
int g_capture = -1;
bool g_replaying = false;
int g_order = -1;                    // 0 fifo, 1 lifo, 2 random
uint64_t g_rng = 0x9E3779B97F4A7C15ull;

void park()
{
    g_stats.switches++;
    ctx_switch(g_blk->fibers[g_cur].ctx, g_sched);
}

void wake(std::vector<int> &list)
{
    for (int f : list) g_blk->ready.push_back(f);
    list.clear();
}

void finish_collective(Warp &w)
{
    std::memcpy(w.res[w.gen & 1], w.pad, sizeof(w.pad));
    w.arrived = 0;
    w.gen++;
    wake(w.waiters);
}

void finish_barrier(Block &b)
{
    b.bar_res[b.bar_gen & 1] = b.bar_pred;
    b.bar_pred = 0;
    b.bar_arrived = 0;
    b.bar_gen++;
    wake(b.bar_waiters);
}

void fiber_main()
{
    (*g_body)();
    Block &b = *g_blk;
    Fiber &f = b.fibers[g_cur];
    f.done = true;
    b.live--;
    Warp &w = b.warps[f.linear >> 5];
    const unsigned lane = f.linear & 31;
    w.live--;
    w.live_mask &= ~(1u << lane);
    w.pad[lane] = 0;
    if (w.arrived > 0 && w.arrived == w.live) { g_stats.collectives_with_exited_lanes++; finish_collective(w); }
    if (b.bar_arrived > 0 && b.bar_arrived == b.live) finish_barrier(b);
    ctx_switch(f.ctx, g_sched);               // never resumed
    abort();
}

void ensure_stacks(size_t n)
{
    if (n <= g_stack_count) return;
    if (g_stacks) munmap(g_stacks, g_stack_count * STACK_BYTES);
    g_stacks = (unsigned char *)mmap(nullptr, n * STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (g_stacks == MAP_FAILED) { perror("emu: mmap of fiber stacks"); abort(); }
    g_stack_count = n;
}

}  // namespace

Stats &stats() { return g_stats; }

unsigned lane_id() { return g_blk->fibers[g_cur].linear & 31; }

unsigned live_lane_mask() { return g_blk->warps[g_blk->fibers[g_cur].linear >> 5].live_mask; }

unsigned char *dyn_smem() { return g_blk->smem; }

int block_barrier(int pred)
{
    Block &b = *g_blk;
    g_stats.barriers++;
    const unsigned gen = b.bar_gen;
    b.bar_pred |= pred ? 1 : 0;
    b.bar_arrived++;
    if (b.bar_arrived == b.live) finish_barrier(b);
    else { b.bar_waiters.push_back(g_cur); park(); }
    return b.bar_res[gen & 1];
}

const uint64_t *warp_gather(uint64_t v)
{
    Block &b = *g_blk;
    const unsigned linear = b.fibers[g_cur].linear;
    Warp &w = b.warps[linear >> 5];
    g_stats.collectives++;
    if (w.live != w.created && w.arrived == 0) g_stats.collectives_with_exited_lanes++;     // legal only with a partial mask
    const unsigned gen = w.gen;
    w.pad[linear & 31] = v;
    w.arrived++;
    if (w.arrived == w.live) finish_collective(w);
    else { w.waiters.push_back(g_cur); park(); }
    return w.res[gen & 1];
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body)
{
    g_stats.launches++;
    if (g_capture >= 0 && !g_replaying) g_graphs[g_capture].push_back(GraphNode{0, grid, block, smem_bytes, body, nullptr, nullptr, 0, 0});
    if (g_order < 0) {
        const char *e = getenv("TKB_EMU_ORDER");
        g_order = !e ? 0 : (!strncmp(e, "lifo", 4) ? 1 : (!strncmp(e, "random", 6) ? 2 : 0));
        if (e && !strncmp(e, "random:", 7)) g_rng ^= strtoull(e + 7, nullptr, 10) * 0x2545F4914F6CDD1Dull;
    }
    const unsigned nthreads = block.x * block.y * block.z;
    if (nthreads == 0 || nthreads > 1024 || grid.x == 0 || grid.y == 0 || grid.z == 0 || smem_bytes > 227 * 1024) {
        g_error = "emu: invalid launch configuration";
        g_last = cudaErrorInvalidValue;
        return;
    }
    ensure_stacks(nthreads);
    std::vector<unsigned char> smem_store(smem_bytes + 64 + sizeof(uint64_t) * 4);
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_store.data() + 15) & ~(uintptr_t)15);
    gridDim = grid;
    blockDim = block;
    g_body = &body;
    for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
    for (unsigned bx = 0; bx < grid.x; bx++) {
        g_stats.blocks++;
        static Block blk;                                           // reused: a fresh 1024-fiber block per CTA costs more than small kernels
        if (blk.fibers.size() < nthreads) blk.fibers.resize(nthreads);
        blk.warps.assign((nthreads + 31) / 32, Warp());
        blk.ready.clear();
        blk.bar_waiters.clear();
        blk.live = (int)nthreads;
        blk.bar_arrived = blk.bar_pred = 0;
        blk.bar_gen = 0;
        blk.smem = smem;
        blk.smem_bytes = smem_bytes;
        std::memset(smem, 0xA5, smem_bytes);                        // shared memory starts uninitialised, not zeroed
        for (int i = 0; i < 4; i++) std::memcpy(smem + smem_bytes + 8 * i, &CANARY, 8);
        g_blk = &blk;
        blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
        for (unsigned t = 0; t < nthreads; t++) {
            Fiber &f = blk.fibers[t];
            f.linear = t;
            f.tid.x = t % block.x; f.tid.y = (t / block.x) % block.y; f.tid.z = t / (block.x * block.y);
            Warp &w = blk.warps[t >> 5];
            w.live++;
            w.created++;
            w.live_mask |= 1u << (t & 31);
            f.done = false;
            ctx_make(f.ctx, g_stacks + (size_t)t * STACK_BYTES, STACK_BYTES, fiber_main);
            blk.ready.push_back((int)t);
            g_stats.fibers++;
        }
        while (!blk.ready.empty()) {
            // Which runnable fiber goes next is unspecified in CUDA (between two synchronisation points the threads of a block
            // run in any order): TKB_EMU_ORDER=fifo (default: ascending thread order), lifo, or random[:seed] -- a kernel whose
            // result depends on the choice has a race.
            if (g_order == 1) {
                g_cur = blk.ready.back();
                blk.ready.pop_back();
            } else if (g_order == 2) {
                g_rng = g_rng * 6364136223846793005ull + 1442695040888963407ull;
                const size_t pick = (size_t)((g_rng >> 33) % blk.ready.size());
                g_cur = blk.ready[pick];
                blk.ready[pick] = blk.ready.back();
                blk.ready.pop_back();
            } else {
                g_cur = blk.ready.front();
                blk.ready.pop_front();
            }
            threadIdx = blk.fibers[g_cur].tid;
            g_stats.switches++;
            ctx_switch(g_sched, blk.fibers[g_cur].ctx);
        }
        g_cur = -1;
        if (blk.live != 0) {                                        // parked fibers nobody will wake: a divergent barrier / collective
            char msg[256];
            int at_bar = (int)blk.bar_waiters.size(), at_warp = 0;
            for (auto &w : blk.warps) at_warp += (int)w.waiters.size();
            snprintf(msg, sizeof msg, "emu: deadlock in block (%u,%u,%u): %d threads alive, %d parked at __syncthreads, %d at a warp collective",
                     bx, by, bz, blk.live, at_bar, at_warp);
            g_error = msg;
            g_last = cudaErrorLaunchFailure;
            g_blk = nullptr;
            return;
        }
        for (int i = 0; i < 4; i++) {
            uint64_t c;
            std::memcpy(&c, smem + smem_bytes + 8 * i, 8);
            if (c != CANARY) {
                g_error = "emu: write beyond the end of dynamic shared memory";
                g_last = cudaErrorLaunchFailure;
                g_blk = nullptr;
                return;
            }
        }
        g_blk = nullptr;
    }
}

}  // namespace emu

extern "C" {
const char *emu_last_error(void) { return emu::g_error.c_str(); }
void emu_stats(long long *o)
{
    const emu::Stats &s = emu::g_stats;
    o[0] = s.launches; o[1] = s.blocks; o[2] = s.fibers; o[3] = s.switches; o[4] = s.collectives; o[5] = s.barriers;
    o[6] = s.collectives_with_exited_lanes;
}
void emu_reset_stats(void) { emu::g_stats = emu::Stats{}; emu::g_error.clear(); }
int emu_graph_begin(void)
{
    emu::g_graphs.emplace_back();
    emu::g_capture = (int)emu::g_graphs.size() - 1;
    return emu::g_capture;
}
int emu_graph_end(void)
{
    const int id = emu::g_capture;
    emu::g_capture = -1;
    return id < 0 ? -1 : (int)emu::g_graphs[id].size();
}
int emu_graph_replay(int id)
{
    if (id < 0 || id >= (int)emu::g_graphs.size() || emu::g_capture >= 0) return -1;
    emu::g_replaying = true;
    for (const emu::GraphNode &n : emu::g_graphs[id]) {
        if (n.kind == 0) emu::launch(n.grid, n.block, n.smem, n.body);
        else if (n.kind == 1) std::memset(n.dst, n.value, n.bytes);
        else std::memmove(n.dst, n.src, n.bytes);
    }
    emu::g_replaying = false;
    return emu::g_last == cudaSuccess ? 0 : 1;
}
}

// ---- runtime API shim ------------------------------------------------------------------------------------------------
cudaError_t cudaGetLastError(void)
{
    const cudaError_t e = emu::g_last;
    emu::g_last = cudaSuccess;
    return e;
}
const char *cudaGetErrorString(cudaError_t e)
{
    if (e == cudaSuccess) return "no error";
    return emu::g_error.empty() ? "emulated CUDA error" : emu::g_error.c_str();
}
// cudaMalloc'ed memory is backed by a POSIX shared-memory object, so that the CUDA IPC calls (the push exchange's peer-mapped
// receive buffers, one process per "GPU") work between emulated processes: the handle carries the object's name, opening it
// maps the same pages at another address -- like a real peer mapping.
namespace {
struct Alloc { std::string name; size_t bytes; bool owner; };
std::map<void *, Alloc> g_allocs;
int g_alloc_seq = 0;
struct UnlinkAtExit {                                               // scratch buffers the library never frees
    ~UnlinkAtExit() { for (auto &kv : g_allocs) if (kv.second.owner && !kv.second.name.empty()) shm_unlink(kv.second.name.c_str()); }
} g_unlink_at_exit;

void *map_shm(const std::string &name, size_t bytes, bool create)
{
    const int fd = shm_open(name.c_str(), create ? (O_CREAT | O_EXCL | O_RDWR) : O_RDWR, 0600);
    if (fd < 0) return nullptr;
    // posix_fallocate reserves the pages now: a full /dev/shm fails here instead of raising SIGBUS at the first touch
    if (create && (ftruncate(fd, (off_t)bytes) != 0 || posix_fallocate(fd, 0, (off_t)bytes) != 0)) {
        close(fd);
        shm_unlink(name.c_str());
        return nullptr;
    }
    void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    return p == MAP_FAILED ? nullptr : p;
}
}  // namespace

cudaError_t cudaMalloc(void **p, size_t bytes)
{
    const size_t padded = (bytes + 4095) / 4096 * 4096 + 4096;
    char name[64];
    snprintf(name, sizeof name, "/tkbemu_%d_%d", (int)getpid(), g_alloc_seq++);
    *p = map_shm(name, padded, true);
    if (!*p) {                                                      // no room in /dev/shm: private memory (no IPC handle for it)
        *p = mmap(nullptr, padded, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (*p == MAP_FAILED) { *p = nullptr; return cudaErrorMemoryAllocation; }
        g_allocs[*p] = Alloc{"", padded, true};
        return cudaSuccess;
    }
    g_allocs[*p] = Alloc{name, padded, true};
    return cudaSuccess;
}
cudaError_t cudaFree(void *p)
{
    if (!p) return cudaSuccess;
    auto it = g_allocs.find(p);
    if (it == g_allocs.end() || !it->second.owner) return cudaErrorInvalidValue;
    munmap(p, it->second.bytes);
    if (!it->second.name.empty()) shm_unlink(it->second.name.c_str());
    g_allocs.erase(it);
    return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind, cudaStream_t)
{
    if (emu::g_capture >= 0 && !emu::g_replaying)
        emu::g_graphs[emu::g_capture].push_back(emu::GraphNode{2, dim3(), dim3(), 0, nullptr, dst, src, 0, bytes});
    std::memmove(dst, src, bytes);
    return cudaSuccess;
}
cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind) { std::memmove(dst, src, bytes); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *dst, int value, size_t bytes, cudaStream_t)
{
    if (emu::g_capture >= 0 && !emu::g_replaying)
        emu::g_graphs[emu::g_capture].push_back(emu::GraphNode{1, dim3(), dim3(), 0, nullptr, dst, nullptr, value, bytes});
    std::memset(dst, value, bytes);
    return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaGetDevice(int *dev) { *dev = 0; return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int *value, cudaDeviceAttr attr, int)
{
    *value = attr == cudaDevAttrMultiProcessorCount ? 148 : 0;
    return cudaSuccess;
}
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p)
{
    auto it = g_allocs.find(p);
    if (it == g_allocs.end() || it->second.name.empty()) return cudaErrorInvalidValue;
    std::memset(h, 0, sizeof *h);
    snprintf(h->reserved, 48, "%s", it->second.name.c_str());
    const uint64_t bytes = it->second.bytes;
    std::memcpy(h->reserved + 48, &bytes, 8);
    return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned)
{
    h.reserved[47] = 0;
    uint64_t bytes = 0;
    std::memcpy(&bytes, h.reserved + 48, 8);
    for (auto &kv : g_allocs)                                       // CUDA refuses to open a handle in the exporting process
        if (kv.second.owner && kv.second.name == h.reserved) return cudaErrorInvalidValue;
    *p = map_shm(h.reserved, bytes, false);
    if (!*p) return cudaErrorInvalidValue;
    g_allocs[*p] = Alloc{h.reserved, bytes, false};
    return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void *p)
{
    auto it = g_allocs.find(p);
    if (it == g_allocs.end() || it->second.owner) return cudaErrorInvalidValue;
    munmap(p, it->second.bytes);
    g_allocs.erase(it);
    return cudaSuccess;
}
