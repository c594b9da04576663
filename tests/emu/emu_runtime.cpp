// TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT (see shim/cuda_runtime.h).
// Kernel launches of the emulated build: CTAs run one after another; the threads of a CTA are ucontext fibers that the
// scheduler resumes in thread order, each until its next __syncthreads() (or its end).  One sweep over the live fibers
// is one barrier phase, so code between two barriers runs thread 0, 1, 2, ... sequentially -- a read of shared memory
// that lacks a barrier after a HIGHER-numbered thread's write sees stale data here, as it may on the device.
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <ucontext.h>

#include <unistd.h>

#include <cstdio>
#include <map>
#include <mutex>
#include <vector>

namespace genfft_emu {

uint3 g_threadIdx, g_blockIdx;
dim3 g_blockDim, g_gridDim;

namespace {
constexpr size_t kStackBytes = 256 * 1024;
constexpr size_t kSmemBytes = 256 * 1024;

struct Fiber {
  ucontext_t ctx;
  void* stack = nullptr;
  bool done = true;
  bool at_barrier = false;  // waiting in __syncthreads(); a fiber that yields while polling (spin_yield) stays runnable
};

std::mutex g_mu;  // one launch at a time (the globals above are process-wide)
std::vector<Fiber> g_fibers;
ucontext_t g_main;
const std::function<void()>* g_body = nullptr;
Fiber* g_current = nullptr;
alignas(128) unsigned char g_smem[kSmemBytes];
unsigned long long g_switches = 0;

void fiber_entry() {
  (*g_body)();
  g_current->done = true;
  // returning resumes uc_link (the scheduler)
}
}  // namespace

unsigned char* dyn_smem() { return g_smem; }

namespace {
std::mutex g_alloc_mu;
std::map<void*, std::pair<void*, size_t>> g_allocs;  // user pointer -> (mapping, length)
}

void* guarded_alloc(size_t bytes) {
  const size_t page = (size_t)sysconf(_SC_PAGESIZE);
  const size_t body = ((bytes ? bytes : 1) + 15) & ~(size_t)15;
  const size_t body_pages = (body + page - 1) / page * page;
  const size_t len = body_pages + 2 * page;
  void* m = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (m == MAP_FAILED) return nullptr;
  unsigned char* base = static_cast<unsigned char*>(m);
  mprotect(base, page, PROT_NONE);
  mprotect(base + page + body_pages, page, PROT_NONE);
  unsigned char* user = base + page + body_pages - body;
  memset(base + page, 0xCD, body_pages);  // poison: reading what nothing wrote shows up as garbage, as on the device
  std::lock_guard<std::mutex> lk(g_alloc_mu);
  g_allocs[user] = {m, len};
  return user;
}

void guarded_free(void* p) {
  if (!p) return;
  std::lock_guard<std::mutex> lk(g_alloc_mu);
  auto it = g_allocs.find(p);
  if (it == g_allocs.end()) {
    fprintf(stderr, "genfft_emu: cudaFree of a pointer cudaMalloc did not return\n");
    abort();
  }
  munmap(it->second.first, it->second.second);
  g_allocs.erase(it);
}

int num_sms() {
  const char* s = getenv("GENFFT_EMU_SMS");
  const int v = s ? atoi(s) : 3;
  return v > 0 ? v : 3;
}

void barrier() {
  g_switches++;
  g_current->at_barrier = true;
  swapcontext(&g_current->ctx, &g_main);
}

// a thread that polls for another thread's action (mbarrier wait) gives the others a turn without arriving at a barrier
void spin_yield() {
  g_switches++;
  swapcontext(&g_current->ctx, &g_main);
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& thread_body) {
  // GENFFT_EMU_NOLAUNCH=1: skip the kernels -- what remains of an exec call is the library's host-side overhead
  static const bool nolaunch = getenv("GENFFT_EMU_NOLAUNCH") != nullptr;
  if (nolaunch) return;
  std::lock_guard<std::mutex> lk(g_mu);
  static const int order = [] {
    const char* o = getenv("GENFFT_EMU_ORDER");
    return !o ? 0 : !strcmp(o, "reverse") ? 1 : !strcmp(o, "shuffle") ? 2 : 0;
  }();
  const size_t nthreads = (size_t)block.x * block.y * block.z;
  if (nthreads == 0 || nthreads > 1024 || smem > kSmemBytes) {
    fprintf(stderr, "genfft_emu: bad launch (%zu threads, %zu bytes of shared memory)\n", nthreads, smem);
    abort();
  }
  if (g_fibers.size() < nthreads) g_fibers.resize(nthreads);
  for (size_t t = 0; t < nthreads; t++) {
    if (!g_fibers[t].stack) {
      void* st = mmap(nullptr, kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
      if (st == MAP_FAILED) abort();
      g_fibers[t].stack = st;
    }
  }
  g_body = &thread_body;
  g_blockDim = block;
  g_gridDim = grid;
  for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
      for (unsigned bx = 0; bx < grid.x; bx++) {
        memset(g_smem, 0xEE, smem);  // shared memory is uninitialised at CTA start
        g_blockIdx = uint3{bx, by, bz};
        for (size_t t = 0; t < nthreads; t++) {
          Fiber& f = g_fibers[t];
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = f.stack;
          f.ctx.uc_stack.ss_size = kStackBytes;
          f.ctx.uc_link = &g_main;
          makecontext(&f.ctx, fiber_entry, 0);
          f.done = false;
          f.at_barrier = false;
        }
        size_t alive = nthreads;
        unsigned long long phase = 0;
        while (alive) {
          // One sweep runs every runnable fiber until it reaches a barrier, polls, or ends.  The barrier opens when
          // every live fiber waits in it.  Order of the threads within a sweep (GENFFT_EMU_ORDER): forward (default),
          // reverse, or a different pseudo-random rotation + direction per sweep -- a missing __syncthreads() between
          // a write and a read shows up as stale data in at least one of them, whichever side has the higher index.
          phase++;
          size_t ran = 0;
          for (size_t slot = 0; slot < nthreads; slot++) {
            size_t t = slot;
            if (order == 1) t = nthreads - 1 - slot;
            else if (order == 2) {
              const unsigned long long h = (phase * 0x9E3779B97F4A7C15ull) ^ ((unsigned long long)bx * 0xD1B54A32D192ED03ull);
              const size_t rot = (size_t)((h >> 17) % nthreads);
              t = (h & 1) ? (rot + slot) % nthreads : (rot + nthreads - slot) % nthreads;
            }
            Fiber& f = g_fibers[t];
            if (f.done || f.at_barrier) continue;
            const unsigned tx = (unsigned)(t % block.x), ty = (unsigned)((t / block.x) % block.y),
                           tz = (unsigned)(t / ((size_t)block.x * block.y));
            g_threadIdx = uint3{tx, ty, tz};
            g_current = &f;
            swapcontext(&g_main, &f.ctx);
            ran++;
            if (f.done) alive--;
          }
          if (ran == 0)  // every live fiber waits in the barrier: open it
            for (size_t t = 0; t < nthreads; t++) g_fibers[t].at_barrier = false;
        }
      }
  g_body = nullptr;
  g_current = nullptr;
}

}  // namespace genfft_emu

extern "C" unsigned long long genfft_emu_fiber_switches() { return genfft_emu::g_switches; }
