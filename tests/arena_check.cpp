// Randomised check of the scratch arena of a cl2 context (cl2_common.cuh: cl2_arena): blocks handed out never overlap,
// nothing leaks, free neighbours are always coalesced.  Built and run by tests/test_arena.py (host only).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <map>
#include <iterator>
#include <cuda_runtime.h>
#include "cl2_common.cuh"
int g_cl2_sync_mode = 0;
int main() {
    cl2_arena a;
    // two chunks that happen to be adjacent in the address space (halves of one buffer) and a third one elsewhere: a
    // block must never span two chunks
    std::vector<char> mem1(1 << 21), mem2(1 << 20);
    const size_t half = mem1.size() / 2;
    char *cut = mem1.data() + half;
    a.add_chunk(mem1.data(), half);
    a.add_chunk(cut, half);
    a.add_chunk(mem2.data(), mem2.size());
    struct B { char *p; size_t n; };
    std::vector<B> live;
    srand(1);
    size_t total = mem1.size() + mem2.size();
    for (int it = 0; it < 200000; it++) {
        if (live.empty() || rand() % 2) {
            size_t n = ((rand() % 64) + 1) * 512;
            char *p = a.take(n);
            if (!p) continue;
            for (auto &b : live) if (p < b.p + b.n && b.p < p + n) { printf("overlap\n"); return 1; }
            bool in1a = p >= mem1.data() && p + n <= cut;
            bool in1b = p >= cut && p + n <= mem1.data() + mem1.size();
            bool in2 = p >= mem2.data() && p + n <= mem2.data() + mem2.size();
            if (!in1a && !in1b && !in2) { printf("block outside a chunk or across two\n"); return 1; }
            live.push_back({p, n});
        } else {
            size_t k = rand() % live.size();
            a.put(live[k].p, live[k].n);
            live[k] = live.back();
            live.pop_back();
        }
        size_t fr = 0;
        if (it % 1000 == 0) {
            for (auto &kv : a.free_blocks) fr += kv.second;
            size_t lv = 0;
            for (auto &b : live) lv += b.n;
            if (fr + lv != total) { printf("leak %zu %zu\n", fr, lv); return 1; }
            // coalesced: no two free blocks adjacent, except across the boundary of two chunks
            char *end = nullptr;
            for (auto &kv : a.free_blocks) { if (end == kv.first && kv.first != cut) { printf("uncoalesced\n"); return 1; } end = kv.first + kv.second; }
        }
    }
    for (auto &b : live) a.put(b.p, b.n);
    if (a.free_blocks.size() != 3) { printf("expected the three chunks back, got %zu blocks\n", a.free_blocks.size()); return 1; }
    printf("ok free blocks %zu\n", a.free_blocks.size());
    return 0;
}
