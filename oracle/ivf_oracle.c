/* CPU oracle (C restatement) for IVFFlat / flat inner-product top-k.  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED: restates the published FAISS IVFFlat-IP search algorithm (faiss-cpu>=1.7.4,
 * requirements.txt:2 of the reference; the library is absent from /root/reference and from this
 * image).  Follows the call sites src/models/faiss_index.py:113 and :145 (index.search) and the
 * semantics listed in SURVEY.md Appendix B:
 *   - quantizer = IndexFlatIP: top-nprobe centroids by inner product, descending;
 *   - the nprobe lists are scanned in that order, every vector scored with a plain fp32 dot;
 *   - a k-min-heap keeps the best k; a candidate enters only with a strictly greater score;
 *   - results are emitted in descending score order; unfilled slots are (-FLT_MAX, -1);
 *   - OpenMP parallel over queries.
 * Tie rule made deterministic: heap order is (score asc, scan-sequence desc) so the final order is
 * (score desc, scan position asc) — identical to oracle/ivf_oracle.py::ivf_search.
 *
 * Used only by tests/ and bench.py's cpu_baseline / --impl reference legs.
 */
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float s; int64_t seq; int64_t id; } ent_t;

/* "a is worse than b": lower score, or same score and scanned later */
static inline int worse(const ent_t* a, const ent_t* b) {
    return a->s < b->s || (a->s == b->s && a->seq > b->seq);
}

static void sift_down(ent_t* h, int n, int i) {
    ent_t v = h[i];
    for (;;) {
        int c = 2 * i + 1;
        if (c >= n) break;
        if (c + 1 < n && worse(&h[c + 1], &h[c])) c++;
        if (!worse(&h[c], &v)) break;
        h[i] = h[c];
        i = c;
    }
    h[i] = v;
}

static inline void heap_offer(ent_t* h, int k, float s, int64_t seq, int64_t id) {
    ent_t e = {s, seq, id};
    if (worse(&h[0], &e)) { h[0] = e; sift_down(h, k, 0); }
}

static int cmp_best_first(const void* pa, const void* pb) {
    const ent_t* a = (const ent_t*)pa; const ent_t* b = (const ent_t*)pb;
    if (worse(b, a)) return -1;
    if (worse(a, b)) return 1;
    return 0;
}

static inline float dotf(const float* a, const float* b, int d) {
    float s = 0.f;
#pragma omp simd reduction(+ : s)
    for (int i = 0; i < d; i++) s += a[i] * b[i];
    return s;
}

static void heap_init(ent_t* h, int k) {
    for (int i = 0; i < k; i++) { h[i].s = -FLT_MAX; h[i].seq = INT64_MAX - i; h[i].id = -1; }
}

static void heap_emit(ent_t* h, int k, float* os, int64_t* oi) {
    qsort(h, (size_t)k, sizeof(ent_t), cmp_best_first);
    for (int i = 0; i < k; i++) { os[i] = h[i].s; oi[i] = h[i].id; }
}

int ivf_oracle_search(const float* q, int64_t nq, int D, const float* centroids, int nlist, int nprobe,
                      const int64_t* offsets, const int64_t* ids, const float* vecs, int k,
                      float* out_scores, int64_t* out_ids, int threads) {
    if (nprobe > nlist) nprobe = nlist;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    int fail = 0;
#pragma omp parallel
    {
        ent_t* heap = (ent_t*)malloc(sizeof(ent_t) * (size_t)(k > nprobe ? k : nprobe));
        ent_t* ph = (ent_t*)malloc(sizeof(ent_t) * (size_t)nprobe);
        if (!heap || !ph) fail = 1;
#pragma omp for schedule(dynamic, 8)
        for (int64_t i = 0; i < nq; i++) {
            if (fail) continue;
            const float* qi = q + i * D;
            /* coarse quantizer: top-nprobe centroids, (score desc, list id asc) */
            heap_init(ph, nprobe);
            for (int c = 0; c < nlist; c++) heap_offer(ph, nprobe, dotf(qi, centroids + (int64_t)c * D, D), c, c);
            qsort(ph, (size_t)nprobe, sizeof(ent_t), cmp_best_first);
            /* list scan */
            heap_init(heap, k);
            int64_t seq = 0;
            for (int p = 0; p < nprobe; p++) {
                int64_t l = ph[p].id;
                if (l < 0) continue;
                for (int64_t j = offsets[l]; j < offsets[l + 1]; j++, seq++)
                    heap_offer(heap, k, dotf(qi, vecs + j * D, D), seq, ids[j]);
            }
            heap_emit(heap, k, out_scores + i * k, out_ids + i * k);
        }
        free(heap); free(ph);
    }
    return fail;
}

int flat_oracle_search(const float* q, int64_t nq, int D, const float* x, int64_t n, int k,
                       float* out_scores, int64_t* out_ids, int threads) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    int fail = 0;
#pragma omp parallel
    {
        ent_t* heap = (ent_t*)malloc(sizeof(ent_t) * (size_t)k);
        if (!heap) fail = 1;
#pragma omp for schedule(dynamic, 1)
        for (int64_t i = 0; i < nq; i++) {
            if (fail) continue;
            const float* qi = q + i * D;
            heap_init(heap, k);
            for (int64_t j = 0; j < n; j++) heap_offer(heap, k, dotf(qi, x + j * D, D), j, j);
            heap_emit(heap, k, out_scores + i * k, out_ids + i * k);
        }
        free(heap);
    }
    return fail;
}
