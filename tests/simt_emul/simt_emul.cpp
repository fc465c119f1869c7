// simt_emul.cpp — TEST-ONLY CPU emulation of the warp primitives in csrc/cz_simt.h.
//
// Each emulated warp is 32 fibers on one OS thread.  A collective (ballot / shfl / sync)
// stores the lane's contribution and switches to the next lane; when control comes back
// every lane has contributed.  Two alternating exchange buffers let fast lanes start the
// next collective while slow lanes still read the previous one.  Blocks run in parallel
// on a small pool of OS threads; warps of a block run one after another (the integer
// kernels never synchronise across warps).
//
// This file is never linked into libcczero_b200.so; it exists so that `pytest -m "not gpu"`
// can execute the device source of the board / tree kernels without a GPU.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <functional>
#include <thread>
#include <vector>
#include <atomic>

namespace czs {

thread_local int tl_lane = 0;
thread_local int tl_warp = 0;
thread_local int tl_nwarps = 1;
thread_local int tl_block = 0;
thread_local unsigned char* tl_smem = nullptr;

extern "C" void cz_fiber_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl cz_fiber_switch
.type cz_fiber_switch,@function
cz_fiber_switch:
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
.size cz_fiber_switch,.-cz_fiber_switch
)");

static const size_t kStack = 256 * 1024;

struct WarpState {
  void* sp[32];
  void* main_sp;
  bool done[32];
  unsigned char* stacks;
  uint64_t xbuf[2][32];
  unsigned bal[2];
  int phase[32];
  const std::function<void()>* body;
};

static thread_local WarpState* tl_ws = nullptr;

static void switch_to_next() {
  WarpState* w = tl_ws;
  const int me = tl_lane;
  for (int k = 1; k <= 32; ++k) {
    const int nx = (me + k) & 31;
    if (!w->done[nx]) {
      if (nx == me) return;
      tl_lane = nx;
      cz_fiber_switch(&w->sp[me], w->sp[nx]);
      tl_lane = me;
      return;
    }
  }
  // everyone else finished
}

static void fiber_main() {
  WarpState* w = tl_ws;
  const int me = tl_lane;
  (*w->body)();
  w->done[me] = true;
  for (int k = 1; k < 32; ++k) {
    const int nx = (me + k) & 31;
    if (!w->done[nx]) {
      tl_lane = nx;
      void* dummy;
      cz_fiber_switch(&dummy, w->sp[nx]);
    }
  }
  void* dummy;
  cz_fiber_switch(&dummy, w->main_sp);
  abort();
}

static void run_warp(const std::function<void()>& body, WarpState* w) {
  memset(w->done, 0, sizeof(w->done));
  memset(w->phase, 0, sizeof(w->phase));
  w->bal[0] = w->bal[1] = 0;
  w->body = &body;
  for (int l = 0; l < 32; ++l) {
    uintptr_t top = (uintptr_t)(w->stacks + (size_t)(l + 1) * kStack);
    top &= ~(uintptr_t)15;
    void** s = (void**)top;
    // layout popped by cz_fiber_switch: r15 r14 r13 r12 rbx rbp ret ; keep (rsp+8)%16==0 at entry
    *--s = nullptr;                 // alignment pad / fake return address slot
    *--s = (void*)&fiber_main;      // ret target
    for (int i = 0; i < 6; ++i) *--s = nullptr;
    w->sp[l] = (void*)s;
  }
  tl_ws = w;
  tl_lane = 0;
  cz_fiber_switch(&w->main_sp, w->sp[0]);
}

// ---- collectives -------------------------------------------------------------------
// Every lane of the warp executes the same sequence of collectives, so a per-lane phase
// counter selects the buffer.
unsigned emul_ballot(bool p) {
  WarpState* w = tl_ws;
  const int me = tl_lane;
  const int ph = w->phase[me] & 1;
  w->phase[me]++;
  if (me == 0) w->bal[ph] = 0;      // lane 0 always arrives first at a new collective
  if (p) w->bal[ph] |= 1u << me;
  switch_to_next();
  return w->bal[ph];
}

uint64_t emul_shfl64(uint64_t v, int src) {
  WarpState* w = tl_ws;
  const int me = tl_lane;
  const int ph = w->phase[me] & 1;
  w->phase[me]++;
  w->xbuf[ph][me] = v;
  switch_to_next();
  return w->xbuf[ph][src & 31];
}

void emul_sync() {
  WarpState* w = tl_ws;
  w->phase[tl_lane]++;
  switch_to_next();
}

// ---- launch ------------------------------------------------------------------------
void emul_launch(int nblocks, int nwarps, size_t smem_bytes, const std::function<void()>& body) {
  if (nblocks <= 0) return;
  int nthreads = (int)std::thread::hardware_concurrency();
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 8) nthreads = 8;
  if (nthreads > nblocks) nthreads = nblocks;
  std::atomic<int> next(0);
  auto worker = [&]() {
    WarpState* w = new WarpState;
    w->stacks = (unsigned char*)malloc(32 * kStack);
    unsigned char* smem = (unsigned char*)calloc(1, smem_bytes + 64);
    for (;;) {
      const int b = next.fetch_add(1);
      if (b >= nblocks) break;
      for (int wi = 0; wi < nwarps; ++wi) {
        tl_block = b; tl_warp = wi; tl_nwarps = nwarps; tl_smem = smem;
        run_warp(body, w);
      }
    }
    free(smem);
    free(w->stacks);
    delete w;
  };
  if (nthreads == 1) { worker(); return; }
  std::vector<std::thread> ts;
  for (int i = 0; i < nthreads; ++i) ts.emplace_back(worker);
  for (auto& t : ts) t.join();
}

}  // namespace czs
