/* Plain-C restatement of the DCLIP retrieval arithmetic — TEST INFRASTRUCTURE ONLY.
 *
 * Follows (paths relative to /root/reference):
 *   eval_scripts/flickr30k_eval.py:249-250  row L2 normalisation (numpy, no eps)
 *   eval_scripts/flickr30k_eval.py:254-266  S = C_hat . I_hat^T  (fp32 products, fp32 sums)
 *   eval_scripts/flickr30k_eval.py:29-42    t2i rank of the ground-truth image
 *   eval_scripts/flickr30k_eval.py:48-64    i2t best rank over an image's captions
 *   training/compute_faiss.py:27,79 + training/image_tokenizer.py:262
 *                                            IndexFlatIP add/search = exact inner product,
 *                                            results by descending score.
 *
 * faiss itself is a third-party dependency that is absent from /root/reference and whose
 * version the reference does not pin (no requirements/lock file): "parity unpinned" at
 * that boundary.  The published algorithm of IndexFlatIP is brute-force inner product +
 * per-query k-selection; that is what is restated here.
 *
 * The reference's BLAS sgemm does not define a summation order, so no implementation can
 * be bit-identical to it on near-ties.  The contract both this oracle and the CUDA path
 * implement is therefore fixed and fully specified:
 *     score(q, g) = acc_d,  acc_0 = 0,  acc_{k+1} = fmaf(q[k], g[k], acc_k)   (fp32, k ascending)
 *     norm(x)     = sqrtf(ss_d), ss_{k+1} = fmaf(x[k], x[k], ss_k);  x_hat[k] = x[k] / norm(x)
 *     order       = descending score, ties by ascending index
 *                   (== np.argsort(-s, kind="stable"); the reference's default argsort
 *                   is unstable on ties)
 * Compile with -ffp-contract=off so only the explicit fmaf calls fuse.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

float dclip_oracle_score(const float* q, const float* g, int64_t d) {
    float acc = 0.0f;
    for (int64_t k = 0; k < d; ++k) acc = fmaf(q[k], g[k], acc);
    return acc;
}

/* x_hat = x / ||x||  row-wise; in place allowed (out == x). */
void dclip_oracle_normalize_rows(const float* x, float* out, int64_t n, int64_t d) {
    for (int64_t i = 0; i < n; ++i) {
        const float* r = x + i * d;
        float ss = 0.0f;
        for (int64_t k = 0; k < d; ++k) ss = fmaf(r[k], r[k], ss);
        float nrm = sqrtf(ss);
        for (int64_t k = 0; k < d; ++k) out[i * d + k] = r[k] / nrm;
    }
}

/* S[nq, ng] = Q . G^T with the contract's summation order. */
void dclip_oracle_scores(const float* q, const float* g, int64_t nq, int64_t ng, int64_t d, float* s) {
    for (int64_t i = 0; i < nq; ++i)
        for (int64_t j = 0; j < ng; ++j) s[i * ng + j] = dclip_oracle_score(q + i * d, g + j * d, d);
}

/* rank[p] of gallery row target[p] for query row qidx[p]:
 *   #{j : s_j > s_t} + #{j < t : s_j == s_t}     (stable descending order) */
void dclip_oracle_ranks(const float* q, const float* g, int64_t ng, int64_t d,
                        const int64_t* qidx, const int64_t* target, int64_t npairs, int32_t* rank) {
    for (int64_t p = 0; p < npairs; ++p) {
        const float* qr = q + qidx[p] * d;
        int64_t t = target[p];
        float st = dclip_oracle_score(qr, g + t * d, d);
        int32_t r = 0;
        for (int64_t j = 0; j < ng; ++j) {
            float sj = dclip_oracle_score(qr, g + j * d, d);
            if (sj > st || (sj == st && j < t)) ++r;
        }
        rank[p] = r;
    }
}

/* Exact inner-product top-k, ordered by (-score, +index).  k <= ng not required:
 * missing slots get score -inf, index -1 (faiss convention). */
void dclip_oracle_topk(const float* q, const float* g, int64_t nq, int64_t ng, int64_t d, int64_t k,
                       float* out_s, int64_t* out_i) {
    float* row = (float*)malloc(sizeof(float) * (size_t)(ng > 0 ? ng : 1));
    for (int64_t i = 0; i < nq; ++i) {
        for (int64_t j = 0; j < ng; ++j) row[j] = dclip_oracle_score(q + i * d, g + j * d, d);
        for (int64_t s = 0; s < k; ++s) { out_s[i * k + s] = -INFINITY; out_i[i * k + s] = -1; }
        for (int64_t j = 0; j < ng; ++j) {
            /* insertion into the sorted k-list; strict > keeps the lower index first on ties */
            int64_t pos = k;
            while (pos > 0 && (out_i[i * k + pos - 1] < 0 || row[j] > out_s[i * k + pos - 1])) --pos;
            if (pos >= k) continue;
            for (int64_t s = k - 1; s > pos; --s) { out_s[i * k + s] = out_s[i * k + s - 1]; out_i[i * k + s] = out_i[i * k + s - 1]; }
            out_s[i * k + pos] = row[j];
            out_i[i * k + pos] = j;
        }
    }
    free(row);
}
