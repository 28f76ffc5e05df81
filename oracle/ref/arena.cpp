// arena.cpp -- TEST INFRASTRUCTURE ONLY (oracle/_ref build).
//
// Monotonic (never-reuse) `operator new` for the verbatim reference build.  The reference's
// DistributeOctTree sorts vector<pair<int,ExtractorNode*>> (/root/reference/src/ORBextractor.cc:948),
// so nodes of equal size are ordered by HEAP ADDRESS; with glibc malloc that order depends on heap
// history.  With a bump allocator address order == creation order, which
// makes the reference deterministic and equal to the canonical rule
// "equal size => the most recently created node is split first".
//
// While a frame is being processed (ref_arena_begin .. ref_arena_end) every operator new is served
// from a thread-local bump arena that is reset at the next ref_arena_begin; outside a frame
// (constructors, persistent members) allocations go to malloc.  operator delete ignores arena
// pointers.  Callers (oracle/ref/ref_capi.cpp) copy results out before the next reset.
#include <cstdlib>
#include <cstdint>
#include <cstdio>
#include <new>

namespace {
struct Arena {
    char* base = nullptr; size_t cap = 0, off = 0; bool active = false;
};
thread_local Arena g_arena;
const size_t kArenaBytes = (size_t)1 << 30;   // virtual reservation; pages are touched lazily
}

// ORB_REF_GLIBC_HEAP=1: leave every allocation to glibc malloc -- the reference exactly as shipped, whose quadtree order then depends
// on the heap's history.  Used only by tools/heap_order_report.py to measure how far that behaviour is from the canonical contract.
static bool glibc_heap() { static const bool v = [] { const char* e = std::getenv("ORB_REF_GLIBC_HEAP"); return e && std::atoi(e) != 0; }(); return v; }

extern "C" void ref_arena_begin() {
    if (glibc_heap()) return;
    Arena& a = g_arena;
    if (!a.base) {
        a.base = (char*)std::malloc(kArenaBytes);
        if (!a.base) { std::fprintf(stderr, "ref arena: out of memory\n"); std::abort(); }
        a.cap = kArenaBytes;
    }
    a.off = 0; a.active = true;
}
extern "C" void ref_arena_end() { g_arena.active = false; }

static inline void* arena_alloc(size_t n) {
    Arena& a = g_arena;
    if (a.active) {
        size_t o = (a.off + 15) & ~(size_t)15;
        if (o + n <= a.cap) { a.off = o + n; return a.base + o; }
        std::fprintf(stderr, "ref arena: exhausted\n"); std::abort();
    }
    void* p = std::malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
static inline void arena_free(void* p) {
    if (!p) return;
    Arena& a = g_arena;
    if (a.base && (char*)p >= a.base && (char*)p < a.base + a.cap) return;   // arena memory is never reused within a frame
    std::free(p);
}
void* operator new(size_t n) { return arena_alloc(n); }
void* operator new[](size_t n) { return arena_alloc(n); }
void operator delete(void* p) noexcept { arena_free(p); }
void operator delete[](void* p) noexcept { arena_free(p); }
void operator delete(void* p, size_t) noexcept { arena_free(p); }
void operator delete[](void* p, size_t) noexcept { arena_free(p); }
