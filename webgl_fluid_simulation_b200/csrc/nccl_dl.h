// nccl_dl.h — NCCL bound at run time with dlopen("libnccl.so.2").
//
// The library is used from two kinds of processes: a Python process in which torch has already
// loaded its bundled libnccl.so.2 (dlopen by soname then returns that same image, so there is
// exactly one NCCL in the process), and a plain C/Node process, where the system
// /usr/lib/x86_64-linux-gnu/libnccl.so.2 is found.  Only the point-to-point subset is declared;
// those prototypes have been ABI-stable across NCCL 2.x.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>

namespace ncdl {

typedef struct ncclComm* ncclComm_t;
struct ncclUniqueId { char internal[128]; };
typedef int ncclResult_t;                 // 0 == ncclSuccess
enum { ncclInt8 = 0, ncclFloat32 = 7 };

struct Api {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    const char* why = "";
};

inline Api& api() {
    static Api a;
    static bool tried = false;
    if (tried) return a;
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (a.handle) break;
    }
    if (!a.handle) { a.why = "dlopen(libnccl.so.2) failed"; return a; }
#define NCDL_SYM(field, name)                                             \
    a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.handle, name)); \
    if (!a.field) { a.why = "missing symbol " name; a.handle = nullptr; return a; }
    NCDL_SYM(GetUniqueId, "ncclGetUniqueId")
    NCDL_SYM(CommInitRank, "ncclCommInitRank")
    NCDL_SYM(CommDestroy, "ncclCommDestroy")
    NCDL_SYM(GroupStart, "ncclGroupStart")
    NCDL_SYM(GroupEnd, "ncclGroupEnd")
    NCDL_SYM(Send, "ncclSend")
    NCDL_SYM(Recv, "ncclRecv")
    NCDL_SYM(GetErrorString, "ncclGetErrorString")
#undef NCDL_SYM
    return a;
}

}  // namespace ncdl
