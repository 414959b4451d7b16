// TEST INFRASTRUCTURE — an in-process "NCCL" for the CUDA emulator: every rank is an OS thread of the same
// process, buffers are ordinary memory, collectives are barriers + copies.  Only the calls and data types
// libcpppd uses are implemented.  Exported with the real NCCL names so that libcpppd's dlsym finds them
// when CPPPD_NCCL_LIB points at libcpppd_emul.so.
#include <condition_variable>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "shim/cuda_runtime.h"
#include "shim/nccl.h"

#ifdef CPPPD_EMUL_FAST_SWITCH
// Fiber switch of the CUDA emulator (declared in shim/cuda_runtime.h; defined once, here): push the callee-saved
// registers and the floating point control words, swap the stack pointer, pop them again.  System V x86-64.
asm(R"(
.text
.globl emul_switch
.type emul_switch,@function
emul_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    subq $8, %rsp
    stmxcsr (%rsp)
    fnstcw 4(%rsp)
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    ldmxcsr (%rsp)
    fldcw 4(%rsp)
    addq $8, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emul_switch,.-emul_switch
)");
#endif

namespace {

struct Barrier {
  std::mutex m;
  std::condition_variable cv;
  int count = 0, n = 0;
  unsigned long long gen = 0;
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    const unsigned long long g = gen;
    if (++count == n) {
      count = 0;
      ++gen;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return gen != g; });
    }
  }
};

struct World {
  int n = 0, joined = 0;
  Barrier bar;
  std::vector<const void *> pub;                      // per rank: published buffer of the current collective
  std::vector<std::map<int, std::pair<const void *, size_t>>> sends;  // [src][dst] of the current group
};

struct Pending { bool send; void *buf; size_t bytes; int peer; };

std::mutex g_registry_mutex;
std::map<long long, World *> g_registry;
long long g_next_id = 1;
thread_local std::vector<Pending> g_group;
thread_local int g_group_depth = 0;
thread_local struct ncclComm *g_current_comm = nullptr;  // the rank thread's communicator (one at a time)

size_t type_bytes(ncclDataType_t t) { return (t == ncclFloat64 || t == ncclUint64) ? 8 : 1; }

}  // namespace

struct ncclComm { World *w; int rank; };

extern "C" {

const char *ncclGetErrorString(ncclResult_t) { return "emulated NCCL error"; }

ncclResult_t ncclGetUniqueId(ncclUniqueId *id) {
  std::lock_guard<std::mutex> lk(g_registry_mutex);
  memset(id, 0, sizeof *id);
  const long long v = g_next_id++;
  memcpy(id->internal, &v, sizeof v);
  return ncclSuccess;
}

ncclResult_t ncclCommInitRank(ncclComm_t *comm, int nranks, ncclUniqueId id, int rank) {
  long long key;
  memcpy(&key, id.internal, sizeof key);
  World *w;
  {
    std::lock_guard<std::mutex> lk(g_registry_mutex);
    World *&slot = g_registry[key];
    if (!slot) {
      slot = new World();
      slot->n = nranks;
      slot->bar.n = nranks;
      slot->pub.assign(nranks, nullptr);
      slot->sends.resize(nranks);
    }
    w = slot;
  }
  *comm = new ncclComm{w, rank};
  g_current_comm = *comm;
  w->bar.wait();  // like the real call: returns once every rank has joined
  return ncclSuccess;
}

ncclResult_t ncclCommDestroy(ncclComm_t comm) { delete comm; return ncclSuccess; }

ncclResult_t ncclAllGather(const void *send, void *recv, size_t count, ncclDataType_t t, ncclComm_t c, cudaStream_t) {
  World *w = c->w;
  const size_t bytes = count * type_bytes(t);
  w->pub[c->rank] = send;
  w->bar.wait();
  std::vector<char> tmp(bytes * w->n);
  for (int r = 0; r < w->n; ++r) memcpy(tmp.data() + bytes * r, w->pub[r], bytes);
  w->bar.wait();  // everybody has read before anybody overwrites (send may alias recv)
  memcpy(recv, tmp.data(), tmp.size());
  w->bar.wait();
  return ncclSuccess;
}

ncclResult_t ncclAllReduce(const void *send, void *recv, size_t count, ncclDataType_t t, ncclRedOp_t op, ncclComm_t c,
                           cudaStream_t) {
  World *w = c->w;
  if (op != ncclSum) return ncclInvalidArgument;
  w->pub[c->rank] = send;
  w->bar.wait();
  std::vector<char> tmp(count * type_bytes(t));
  if (t == ncclFloat64) {
    double *out = reinterpret_cast<double *>(tmp.data());
    for (size_t i = 0; i < count; ++i) {
      double acc = static_cast<const double *>(w->pub[0])[i];
      for (int r = 1; r < w->n; ++r) acc += static_cast<const double *>(w->pub[r])[i];
      out[i] = acc;
    }
  } else if (t == ncclUint64) {
    unsigned long long *out = reinterpret_cast<unsigned long long *>(tmp.data());
    for (size_t i = 0; i < count; ++i) {
      unsigned long long acc = 0;
      for (int r = 0; r < w->n; ++r) acc += static_cast<const unsigned long long *>(w->pub[r])[i];
      out[i] = acc;
    }
  } else {
    for (size_t i = 0; i < count; ++i) {
      char acc = 0;
      for (int r = 0; r < w->n; ++r) acc = (char)(acc + static_cast<const char *>(w->pub[r])[i]);
      tmp[i] = acc;
    }
  }
  w->bar.wait();
  memcpy(recv, tmp.data(), tmp.size());
  w->bar.wait();
  return ncclSuccess;
}

ncclResult_t ncclGroupStart() { ++g_group_depth; return ncclSuccess; }

ncclResult_t ncclSend(const void *buf, size_t count, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t) {
  g_current_comm = c;
  g_group.push_back(Pending{true, const_cast<void *>(buf), count * type_bytes(t), peer});
  return ncclSuccess;
}

ncclResult_t ncclRecv(void *buf, size_t count, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t) {
  g_current_comm = c;
  g_group.push_back(Pending{false, buf, count * type_bytes(t), peer});
  return ncclSuccess;
}

// libcpppd always brackets its sends / receives in one group per exchange, and every rank runs the same
// number of exchanges, so a group end is a collective: publish the sends, barrier, copy, barrier.
ncclResult_t ncclGroupEnd() {
  if (--g_group_depth > 0) return ncclSuccess;
  ncclComm_t c = g_current_comm;
  if (!c) return ncclInvalidArgument;
  World *w = c->w;
  for (const Pending &p : g_group)
    if (p.send) w->sends[c->rank][p.peer] = {p.buf, p.bytes};
  w->bar.wait();
  for (const Pending &p : g_group) {
    if (p.send) continue;
    auto it = w->sends[p.peer].find(c->rank);
    if (it == w->sends[p.peer].end() || it->second.second != p.bytes) return ncclInvalidArgument;
    memcpy(p.buf, it->second.first, p.bytes);
  }
  w->bar.wait();
  w->sends[c->rank].clear();
  g_group.clear();
  w->bar.wait();
  return ncclSuccess;
}

}  // extern "C"
