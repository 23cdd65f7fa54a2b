// LRU sector-cache simulator for the x gathers of a row-ordered SpMV.
// usage: lru <trace.bin> <cache_sectors> <elems_per_sector>
// trace.bin: int64 nnz, then int32 col[nnz] in processing order.  Prints misses.
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
typedef struct { int32_t prev, next; int64_t key; } Node;
int main(int argc, char **argv) {
    if (argc < 4) return 2;
    FILE *f = fopen(argv[1], "rb"); if (!f) return 2;
    int64_t nnz; if (fread(&nnz, 8, 1, f) != 1) return 2;
    int32_t *col = malloc(sizeof(int32_t) * nnz);
    if (fread(col, 4, nnz, f) != (size_t)nnz) return 2;
    fclose(f);
    int64_t cap = atoll(argv[2]); int eps = atoi(argv[3]);
    int64_t hsize = 1; while (hsize < cap * 4) hsize <<= 1;
    int32_t *table = malloc(sizeof(int32_t) * hsize); memset(table, 0xff, sizeof(int32_t) * hsize);
    Node *nodes = malloc(sizeof(Node) * cap);
    int32_t head = -1, tail = -1; int64_t used = 0, misses = 0;
    for (int64_t t = 0; t < nnz; t++) {
        int64_t key = col[t] / eps;
        int64_t h = (key * 0x9E3779B97F4A7C15ull) & (hsize - 1);
        int32_t idx = -1;
        while (table[h] != -1) { if (nodes[table[h]].key == key) { idx = table[h]; break; } h = (h + 1) & (hsize - 1); }
        if (idx >= 0) {                       // hit: move to head
            if (idx != head) {
                Node *n = &nodes[idx];
                if (n->prev >= 0) nodes[n->prev].next = n->next;
                if (n->next >= 0) nodes[n->next].prev = n->prev;
                if (idx == tail) tail = n->prev;
                n->prev = -1; n->next = head; nodes[head].prev = idx; head = idx;
            }
            continue;
        }
        misses++;
        int32_t slot;
        if (used < cap) slot = (int32_t)used++;
        else {                                // evict tail: remove from hash (backward-shift deletion)
            slot = tail;
            int64_t ek = nodes[slot].key;
            int64_t eh = (ek * 0x9E3779B97F4A7C15ull) & (hsize - 1);
            while (table[eh] != slot) eh = (eh + 1) & (hsize - 1);
            int64_t hole = eh, nx = (eh + 1) & (hsize - 1);
            table[hole] = -1;
            while (table[nx] != -1) {
                int64_t home = (nodes[table[nx]].key * 0x9E3779B97F4A7C15ull) & (hsize - 1);
                int64_t dist_home = (nx - home) & (hsize - 1), dist_hole = (nx - hole) & (hsize - 1);
                if (dist_home >= dist_hole) { table[hole] = table[nx]; table[nx] = -1; hole = nx; }
                nx = (nx + 1) & (hsize - 1);
            }
            tail = nodes[slot].prev; if (tail >= 0) nodes[tail].next = -1; else head = -1;
        }
        nodes[slot].key = key; nodes[slot].prev = -1; nodes[slot].next = head;
        if (head >= 0) nodes[head].prev = slot; head = slot; if (tail < 0) tail = slot;
        int64_t ih = (key * 0x9E3779B97F4A7C15ull) & (hsize - 1);
        while (table[ih] != -1) ih = (ih + 1) & (hsize - 1);
        table[ih] = slot;
    }
    printf("%lld %lld\n", (long long)nnz, (long long)misses);
    return 0;
}
