// See fiber.hpp.  x86-64 System V only (the library's host side targets the GPU box's Xeon); the
// context switch saves the callee-saved registers on the outgoing stack and swaps stack pointers.
#include "fiber.hpp"
#include "transcript.hpp"

#include <cstdlib>
#include <cstring>
#include <memory>

extern "C" {
int cdl_keccak_x8_available();
void cdl_keccak_f1600_x8(uint64_t* const* st, int n);
void cdl_ctx_switch(void** save_sp, void* load_sp);
}

#if defined(__x86_64__)
asm(R"(
.text
.globl cdl_ctx_switch
.type cdl_ctx_switch,@function
cdl_ctx_switch:
  endbr64
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
.size cdl_ctx_switch,.-cdl_ctx_switch
)");
#define CDL_HAVE_FIBERS 1
#else
#define CDL_HAVE_FIBERS 0
#endif

namespace cdlh {


namespace {

constexpr int kGroup = 8;
constexpr size_t kStackBytes = 256 * 1024;

// Fibers of a group hand the thread to each other directly: a fiber that needs a permutation parks
// its state (PENDING) and switches to the next READY fiber; when nobody is READY every live fiber is
// PENDING, so the parking fiber permutes all parked states with one eight-way call and simply goes on
// (one context switch per permutation and fiber, none for the fiber that triggers the call).
enum : uint8_t { kReady = 0, kRunning, kPending, kDone };

struct Group {
  const std::function<void(size_t)>* fn = nullptr;
  size_t first = 0;
  int count = 0;
  void* main_sp = nullptr;
  void* sp[kGroup] = {};
  uint64_t* pend[kGroup] = {};
  uint8_t state[kGroup] = {};
  int cur = -1;

  int find_ready() const {
    for (int d = 1; d <= count; d++) {
      int k = (cur + d) % count;
      if (state[k] == kReady) return k;
    }
    return -1;
  }
  bool permute_pending() {
    uint64_t* ptrs[kGroup];
    int n = 0;
    for (int k = 0; k < count; k++)
      if (state[k] == kPending) { ptrs[n++] = pend[k]; state[k] = kReady; }
    if (n == 1) keccak_f1600_plain(ptrs[0]);
    else if (n > 1) cdl_keccak_f1600_x8(ptrs, n);
    return n > 0;
  }
};

thread_local Group* tls_group = nullptr;

struct Stacks {
  uint8_t* mem = nullptr;
  Stacks() { mem = static_cast<uint8_t*>(std::aligned_alloc(64, kGroup * kStackBytes)); }
  ~Stacks() { std::free(mem); }
};
thread_local std::unique_ptr<Stacks> tls_stacks;

#if CDL_HAVE_FIBERS
void fiber_entry() {
  Group* g = tls_group;
  const int me = g->cur;
  (*g->fn)(g->first + (size_t)me);
  g->state[me] = kDone;
  for (;;) {  // hand over; never resumed once done
    int nxt = g->find_ready();
    if (nxt < 0 && g->permute_pending()) nxt = g->find_ready();
    if (nxt >= 0) {
      g->cur = nxt;
      g->state[nxt] = kRunning;
      cdl_ctx_switch(&g->sp[me], g->sp[nxt]);
    } else {
      cdl_ctx_switch(&g->sp[me], g->main_sp);
    }
  }
}
#endif

}  // namespace

bool fibers_available() {
#if CDL_HAVE_FIBERS
  static const bool ok = cdl_keccak_x8_available() && !std::getenv("CDL_NO_FIBERS");
  return ok;
#else
  return false;
#endif
}

bool fiber_keccak(uint64_t* st) {
#if CDL_HAVE_FIBERS
  Group* g = tls_group;
  if (!g) return false;
  const int me = g->cur;
  g->pend[me] = st;
  g->state[me] = kPending;
  const int nxt = g->find_ready();
  if (nxt < 0) {  // everybody is parked: permute them all and carry on
    g->permute_pending();
    g->state[me] = kRunning;
    return true;
  }
  g->cur = nxt;
  g->state[nxt] = kRunning;
  cdl_ctx_switch(&g->sp[me], g->sp[nxt]);
  return true;  // resumed by a sibling: cur == me, state == kRunning, st permuted
#else
  (void)st;
  return false;
#endif
}

void run_fiber_group(const std::function<void(size_t)>& fn, size_t first, size_t count) {
#if CDL_HAVE_FIBERS
  if (count > (size_t)kGroup) count = kGroup;
  if (!fibers_available() || count < 2 || tls_group) {  // no nesting: a fiber that calls par() runs it inline
    for (size_t i = 0; i < count; i++) fn(first + i);
    return;
  }
  if (!tls_stacks) tls_stacks.reset(new Stacks());
  Group g;
  g.fn = &fn;
  g.first = first;
  g.count = (int)count;
  for (size_t k = 0; k < count; k++) {
    // ret slot at a 16-byte aligned address: rsp = 8 (mod 16) on entry to fiber_entry, as after a call
    uint8_t* top = tls_stacks->mem + (k + 1) * kStackBytes;
    void** ret = reinterpret_cast<void**>(top - 32);
    ret[0] = reinterpret_cast<void*>(&fiber_entry);
    ret[1] = nullptr;  // fake return address of fiber_entry (never used)
    void** regs = ret - 6;
    for (int i = 0; i < 6; i++) regs[i] = nullptr;
    g.sp[k] = regs;
    g.state[k] = kReady;
  }
  tls_group = &g;
  g.cur = 0;
  g.state[0] = kRunning;
  cdl_ctx_switch(&g.main_sp, g.sp[0]);  // returns when every fiber is done
  tls_group = nullptr;
#else
  for (size_t i = 0; i < count; i++) fn(first + i);
#endif
}

}  // namespace cdlh
