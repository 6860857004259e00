// SMP_beta_b200.h -- the second-order CCN model SMP_beta on the B200 path, in the reference's host language and with the
// reference's own model API (GraphFlow/SMP_beta.h:30-1010; GPU twin GraphFlow_gpu/SMP_beta_gpu.h:199-213, 584-610):
//
//     ccn_b200::SMP_beta net(max_nVertices, nLevels, nChanels, nFeatures, nDepth);
//     net.BatchLearn(nBatch, graphs, targets, learning_rate);   // -> pair(loss before, loss after)   SMP_beta.h:745-772
//     net.Predict(graph);  net.Feature(graph);  net.getLoss(nBatch, graphs, targets);              // :871-879, 931-943, 642-649
//     net.save_model(file);  net.load_model(file);                                                 // :980-1002
//
// With `#define CCN_B200_DROP_IN` before the include the class is also visible as `SMP_beta`, so tests/test_SMP_beta.cpp
// compiles with its one `#include "../GraphFlow/SMP_beta.h"` swapped for this header.
//
// What runs where.  The reference wires ~14 host operators per vertex per level and runs them one graph at a time
// (complete_computation_graph, :531-639).  Here a whole mini-batch of graphs is ONE launch set per level on the device:
//   graph tables        ccn_graph_tables_* (native host code; cached per DenseGraph): Floyd-Warshall, WL features, the
//                       reference's ranking, receptive fields, reduced adjacency, promotion index tables    (:343-529)
//   level 0             f_0[v] = LeakyReLU(H x_v) for all vertices: the mix kernels with a zero bias            (:563-573)
//   level l = 1..L      ccn_gather_level_forward / _backward: promotion + stack + RisiContraction_18 + MatMul(K_l) + b_l +
//                       LeakyReLU, device resident from level to level (Z of level l-1 IS the f buffer of level l)  (:576-618)
//   read-out + loss     ccn_readout_forward / _backward                                                         (:620-639)
//   optimizer           the reference's own Adam object on the host parameters (Adam.h), fed the batch-summed gradients
//                       exactly like SumGradients + sgd->Learn(learning_rate, nBatch)                         (:751-771)
// Only the parameters (up), their gradients, the losses and predictions (down) cross PCIe per call.
// Parameters live in the reference's own host types (Matrix / Vector; same registration order H, K_l, b_l, ..., W) and are
// initialised exactly like weights_initialization (:319-323): the same rand() sequence gives the same model.
//
// Include with one of the reference trees on the include path (-I<GraphFlow>/GraphFlow or -I<GraphFlow>/GraphFlow_32bit); only
// DenseGraph.h, Matrix.h, Vector.h and Adam.h of the reference are used (GraphFlow_32bit/GraphFlow.h does not compile with g++).
// There is NO CPU path: without a usable device the first call aborts with the C-ABI error.
#ifndef GRAPHFLOW_B200_SMP_BETA_B200_H_INCLUDED
#define GRAPHFLOW_B200_SMP_BETA_B200_H_INCLUDED

#include <cassert>
#include <cmath>
#include <cstdint>
#include <fstream>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "Matrix.h"
#include "DenseGraph.h"
#include "Adam.h"
#include "Momentum.h"

#include "ccn_ops_b200.h"

namespace ccn_b200 {

// A grow-only device allocation (bytes).
struct RawDevice {
    void *p;
    size_t cap;
    RawDevice() : p(NULL), cap(0) {}
    void *reserve(size_t bytes) {
        if (bytes > cap) {
            ccn_ctx *ctx = context();
            if (p) {
                CCN_B200_CHECK(ctx, ccn_stream_synchronize(ctx, NULL));
                CCN_B200_CHECK(ctx, ccn_device_free(ctx, p));
            }
            CCN_B200_CHECK(ctx, ccn_device_alloc(ctx, &p, bytes + bytes / 4));
            cap = bytes + bytes / 4;
        }
        return p;
    }
    template <class T>
    T *upload(const std::vector<T> &h) {
        reserve(h.size() * sizeof(T) + 16);
        ccn_ctx *ctx = context();
        if (!h.empty()) CCN_B200_CHECK(ctx, ccn_h2d(ctx, p, &h[0], h.size() * sizeof(T), NULL));
        return static_cast<T *>(p);
    }
    float *floats(size_t n) { return static_cast<float *>(reserve(n * sizeof(float) + 16)); }
    void release() {
        if (p) {
            ccn_ctx *ctx = context();
            ccn_stream_synchronize(ctx, NULL);
            ccn_device_free(ctx, p);
        }
        p = NULL;
        cap = 0;
    }
};

// The model body, shared by the façades below.  Optimizer = the reference's own optimizer class (Adam for SMP_beta / SMP_omega,
// Momentum for SMP_2D_ver8); K_TRANSPOSED = the level weights are stored [C, 18 C] (SMP_2D_ver8's CustomMatMulTensor,
// SMP_2D_ver8.h:130, 526-527) instead of [18 C, C].
template <class Optimizer, bool K_TRANSPOSED>
class SMP_model {
public:
    struct LevelParams {  // the reference's `level[l] -> K`, `level[l] -> b` (SMP_beta.h:195-196)
        Matrix *K;
        Vector *b;
    };

    SMP_model(int max_nVertices, int nLevels, int nChanels, int nFeatures, int nDepth, Optimizer *optimizer) {
        this->max_nVertices = max_nVertices;
        this->nLevels = nLevels;
        this->nChanels = nChanels;
        this->nFeatures = nFeatures;
        this->nDepth = nDepth;
        graph_kind = CCN_GRAPH_BETA;
        max_receptive_field = 0;
        chunk_graphs = 256;
        pass_tables_valid = false;
        head = HEAD_INNER_PRODUCT;
        W1 = NULL;
        W2 = NULL;
        width.assign(nLevels + 1, nChanels);
        H = new Matrix(nChanels, nFeatures * (nDepth + 1));                      // :134
        level = new LevelParams *[nLevels + 1];
        level[0] = NULL;
        for (int l = 1; l <= nLevels; ++l) {
            level[l] = new LevelParams();
            level[l]->K = K_TRANSPOSED ? new Matrix(nChanels, nContractions * nChanels)   // SMP_2D_ver8.h:130
                                       : new Matrix(nContractions * nChanels, nChanels);  // SMP_beta.h:195
            level[l]->b = new Vector(nChanels);                                  // :196
        }
        W = new Vector(nChanels);
        sgd = optimizer;                                                         // :274-280: H, (K_l, b_l)..., W
        sgd->add(H);
        for (int l = 1; l <= nLevels; ++l) {
            sgd->add(level[l]->K);
            sgd->add(level[l]->b);
        }
        sgd->add(W);
        weights_initialization();
    }

    // SMP_omega_physics (SMP_omega_physics.h:31-47, 99-287): raw vertex features (H is [C, nFeatures]), receptive fields limited to
    // max_receptive_field members (:367-418), the channel width halves per level (:142-146), EVERY level feeds the read-out
    // (ShrinkTensor -> LeakyReLU -> SumVectors per level, concatenated, :560-583), which ends in one hidden layer
    // (MatVecMul(W1) -> LeakyReLU -> InnerProduct(W2), :585-595).  Registration order H, (K_l, b_l)..., W1, W2 (:255-262).
    struct PhysicsTag {};
    SMP_model(PhysicsTag, int max_nVertices, int max_receptive_field, int nLevels, int nChanels, int nFeatures, Optimizer *optimizer,
              bool with_head = true) {
        this->max_nVertices = max_nVertices;
        this->nLevels = nLevels;
        this->nChanels = nChanels;
        this->nFeatures = nFeatures;
        this->nDepth = 0;
        graph_kind = CCN_GRAPH_OMEGA;
        this->max_receptive_field = max_receptive_field;
        chunk_graphs = 256;
        pass_tables_valid = false;
        head = HEAD_HIDDEN_LAYER;
        W = NULL;
        width.assign(nLevels + 1, nChanels);
        for (int l = 1; l <= nLevels; ++l) width[l] = std::max(1, width[l - 1] / 2);
        H = new Matrix(nChanels, nFeatures);                                     // :103
        level = new LevelParams *[nLevels + 1];
        level[0] = NULL;
        int total = width[0];
        for (int l = 1; l <= nLevels; ++l) {
            level[l] = new LevelParams();
            level[l]->K = new Matrix(nContractions * width[l - 1], width[l]);
            level[l]->b = new Vector(width[l]);
            total += width[l];
        }
        sgd = optimizer;
        if (!with_head) {  // a trunk of a multi-trunk model (SMP_omega_pairgraphs): the owner registers and initialises the parameters
            W1 = NULL;
            W2 = NULL;
            return;
        }
        W1 = new Matrix(total / 2, total);                                       // :232-234
        W2 = new Vector(total / 2);
        sgd->add(H);
        for (int l = 1; l <= nLevels; ++l) {
            sgd->add(level[l]->K);
            sgd->add(level[l]->b);
        }
        sgd->add(W1);
        sgd->add(W2);
        weights_initialization();
    }

    // width of the graph feature: nChanels (the last level's), or the sum of all level widths for the multi-level read-out
    int feature_width() const {
        if (head == HEAD_INNER_PRODUCT) return nChanels;
        int t = 0;
        for (size_t l = 0; l < width.size(); ++l) t += width[l];
        return t;
    }

    // weights_initialization (:319-323) calls GraphFlow::uniform_init on every parameter through its static type Vector*
    // (GraphFlow.h:1297-1306), for matrices too: value = (rand() % 10) / (10 size), negated when the next rand() is odd.
    void weights_initialization() {
        for (size_t i = 0; i < sgd->params.size(); ++i) {
            Vector *V = sgd->params[i];
            for (int j = 0; j < V->size; ++j) {
                V->value[j] = (double)(rand() % 10) / (10.0 * V->size);
                if (rand() % 2 == 1) V->value[j] = -V->value[j];
            }
        }
    }

    // ---- the reference's calls --------------------------------------------------------------------------------------
    template <class TargetT>
    std::pair<double, double> BatchLearn(int nBatch, DenseGraph **molecule, TargetT *target, double learning_rate) {
        assert(nBatch > 0);
        std::pair<double, double> ret;
        ret.first = run(nBatch, molecule, target, true);       // the loss of this forward IS getLoss at the current weights
        sgd->Learn(learning_rate, nBatch);                     // param->gradient[] hold the sums over the batch (:767-768)
        ret.second = run(nBatch, molecule, target, false);
        return ret;
    }

    template <class TargetT>
    double getLoss(int nBatch, DenseGraph **molecule, TargetT *target) {
        return run(nBatch, molecule, target, false);
    }

    // Forward + backward only: every parameter's gradient[] receives the sum over the batch (what SumGradients holds before
    // sgd->Learn in the reference, :753-767).  Returns the summed loss.
    template <class TargetT>
    double Gradients(int nBatch, DenseGraph **molecule, TargetT *target) {
        return run(nBatch, molecule, target, true);
    }

    double Predict(DenseGraph *molecule) {
        run(1, &molecule, static_cast<double *>(NULL), false);
        return last_predict[0];
    }

    void Predict(int nBatch, DenseGraph **molecule, double *predict) {  // the batched form of Threaded_Predict (:891-929)
        run(nBatch, molecule, static_cast<double *>(NULL), false);
        for (int i = 0; i < nBatch; ++i) predict[i] = last_predict[i];
    }

    std::vector<double> Feature(DenseGraph *molecule) {
        run(1, &molecule, static_cast<double *>(NULL), false);
        return std::vector<double>(last_feature.begin(), last_feature.begin() + feature_width());
    }

    void save_model(std::string filename) {  // :980-990, the same text format
        std::ofstream file(filename.c_str(), std::ios::out);
        for (size_t i = 0; i < sgd->params.size(); ++i)
            for (int j = 0; j < sgd->params[i]->size; ++j) file << sgd->params[i]->value[j] << " ";
        file.close();
    }

    void load_model(std::string filename) {  // :992-1002
        std::ifstream file(filename.c_str(), std::ios::in);
        for (size_t i = 0; i < sgd->params.size(); ++i)
            for (int j = 0; j < sgd->params[i]->size; ++j) file >> sgd->params[i]->value[j];
        file.close();
    }

    // Graph tables are cached per DenseGraph object and re-derived when its adjacency or features change.
    void clear_cache() {
        for (typename std::map<DenseGraph *, Cached>::iterator it = cache.begin(); it != cache.end(); ++it) ccn_graph_tables_destroy(it->second.tables);
        cache.clear();
        pass_tables_valid = false;
    }

    void release() {
        clear_cache();
        RawDevice *all[] = {&d_feat, &d_Ht, &d_zero, &d_pre0, &d_gHt, &d_W, &d_gW, &d_target, &d_shr, &d_gfeat, &d_pred, &d_loss,
                            &d_instptr, &d_instgraph, &d_gX, &d_T, &d_dgf, &d_gact0, &d_zero2};
        for (size_t i = 0; i < sizeof(all) / sizeof(all[0]); ++i) all[i]->release();
        for (size_t l = 0; l < lv.size(); ++l) lv[l].release();
        d_act0.release();
    }

    // Kernel launches of this thread's context so far (evidence that the device path ran).
    long long kernel_launches() const { return (long long)ccn_ctx_kernel_launches(context()); }

    int max_nVertices, nLevels, nChanels, nFeatures, nDepth;
    int graph_kind;           // CCN_GRAPH_BETA, or CCN_GRAPH_OMEGA_WL for SMP_omega
    int max_receptive_field;  // SMP_omega only (0 = unlimited)
    int chunk_graphs;  // graphs per device pass (gradients are accumulated over the passes of one call)
    enum { HEAD_INNER_PRODUCT = 0, HEAD_HIDDEN_LAYER = 1 };
    int head;                 // the read-out head: InnerProduct(W) on the last level's feature, or the multi-level hidden-layer head
    std::vector<int> width;   // channels of level 0..nLevels
    Matrix *H;
    LevelParams **level;
    Vector *W;                // HEAD_INNER_PRODUCT
    Matrix *W1;               // HEAD_HIDDEN_LAYER
    Vector *W2;
    Optimizer *sgd;
    std::vector<double> last_predict, last_feature, last_loss;
    static const int nContractions = 18;

protected:
    struct Cached {
        ccn_graph_tables *tables;
        unsigned long long digest;
    };
    struct LevelDevice {
        RawDevice f_off, m, pos, adj, n, X, Y, Z, gZ, K, b, gK, gb, shr;
        void release() {
            RawDevice *all[] = {&f_off, &m, &pos, &adj, &n, &X, &Y, &Z, &gZ, &K, &b, &gK, &gb, &shr};
            for (size_t i = 0; i < sizeof(all) / sizeof(all[0]); ++i) all[i]->release();
        }
    };

    static unsigned long long digest_of(const DenseGraph *g) {
        unsigned long long h = 1469598103934665603ull;
        for (int i = 0; i < g->nVertices; ++i) {
            for (int j = 0; j < g->nVertices; ++j) h = (h ^ (unsigned long long)(g->adj[i][j] + 7)) * 1099511628211ull;
            for (int f = 0; f < g->nFeatures; ++f) {
                const double x = g->feature[i][f];
                unsigned long long bits;
                std::memcpy(&bits, &x, sizeof(bits));
                h = (h ^ bits) * 1099511628211ull;
            }
        }
        return h ^ (unsigned long long)g->nVertices;
    }

    ccn_graph_tables *tables_of(DenseGraph *g) {
        assert(g->nFeatures == nFeatures);
        assert(g->nVertices <= max_nVertices);
        const unsigned long long d = digest_of(g);
        typename std::map<DenseGraph *, Cached>::iterator it = cache.find(g);
        if (it != cache.end()) {
            if (it->second.digest == d) return it->second.tables;
            ccn_graph_tables_destroy(it->second.tables);
            cache.erase(it);
        }
        const int V = g->nVertices;
        std::vector<int32_t> adj((size_t)V * V);
        std::vector<double> feat((size_t)V * nFeatures);
        for (int i = 0; i < V; ++i) {
            for (int j = 0; j < V; ++j) adj[(size_t)i * V + j] = g->adj[i][j];
            for (int f = 0; f < nFeatures; ++f) feat[(size_t)i * nFeatures + f] = g->feature[i][f];
        }
        Cached c;
        c.digest = d;
        int rc = ccn_graph_tables_create(&adj[0], &feat[0], V, nFeatures, nLevels, nDepth, graph_kind, max_receptive_field, &c.tables);
        if (rc != CCN_OK) die(NULL, rc, "ccn_graph_tables_create");
        cache[g] = c;
        return c.tables;
    }

    // forward (+ backward) of `nBatch` graphs in passes of chunk_graphs; returns the summed loss (0 without targets); with
    // need_grads the parameters' gradient[] arrays receive the sums over the batch.
    template <class TargetT>
    double run(int nBatch, DenseGraph **molecule, TargetT *target, bool need_grads) {
        last_predict.assign(nBatch, 0.0);
        last_loss.assign(nBatch, 0.0);
        last_feature.assign((size_t)nBatch * feature_width(), 0.0);
        if (need_grads)
            for (size_t i = 0; i < sgd->params.size(); ++i)
                for (int j = 0; j < sgd->params[i]->size; ++j) sgd->params[i]->gradient[j] = 0.0;
        double total = 0.0;
        for (int g0 = 0; g0 < nBatch; g0 += chunk_graphs) {
            const int cnt = std::min(chunk_graphs, nBatch - g0);
            total += pass(cnt, molecule + g0, target ? target + g0 : static_cast<TargetT *>(NULL), need_grads, g0);
        }
        return total;
    }

public:
    // ---- one device pass over G graphs, in four steps so that a model made of several trunks (SMP_omega_pairgraphs) can drive them:
    //   levels_forward        graph tables, parameters up, level 0 and the L contraction levels
    //   level_features        (multi-level read-out) every level's ShrinkTensor -> LeakyReLU -> SumVectors into [G, feature_width()]
    //   level_features_grad   the gradient of that matrix up; writes the last level's share of the activation gradient
    //   levels_backward       the levels and level 0 backwards; adds into H / K_l / b_l -> gradient[]
    struct PassState {
        int G;
        int64_t Vtot;
        std::vector<int> n_max;
        std::vector<int64_t> stride;
        float *dFeat, *dHt, *dZero, *pre0, *act0, *shr0, *dDgf, *g_cur;
        const float *f_last;
        int64_t *dPtr;
        int32_t *dIG;
    };
    PassState ps;
    bool pass_tables_valid;
    unsigned long long pass_tables_key;

    void levels_forward(int G, DenseGraph **molecule) {
        ccn_ctx *ctx = context();
        const int C = nChanels, L = nLevels, Fw = nFeatures * (nDepth + 1);  // C = width[0]
        const std::vector<int> &w = width;
        const float alpha = 0.01f;
        // ---- host tables of the pass ----
        std::vector<ccn_graph_tables *> gt(G);
        std::vector<int64_t> vbase(G + 1, 0);
        unsigned long long key = 1469598103934665603ull ^ (unsigned long long)G;  // the graphs of the pass: identity + content, in order
        for (int g = 0; g < G; ++g) {
            gt[g] = tables_of(molecule[g]);
            vbase[g + 1] = vbase[g] + molecule[g]->nVertices;
            key = (key ^ (unsigned long long)(uintptr_t)molecule[g]) * 1099511628211ull;
            key = (key ^ cache[molecule[g]].digest) * 1099511628211ull;
        }
        const int64_t Vtot = vbase[G];
        // The assembled per-level tables of the LAST pass stay on the device: a pass over the same graphs (BatchLearn's second
        // forward, the next epoch over the same mini-batch) skips the host assembly and the uploads.
        const bool reuse = pass_tables_valid && key == pass_tables_key && G == ps.G && Vtot == ps.Vtot;
        std::vector<int> n_max(L + 1, 1);
        std::vector<int64_t> stride(L + 1, C);  // element stride between consecutive vertices' tensors at level l (w[l] n_max[l]^2)
        float *dFeat = NULL;
        if (reuse) {
            n_max = ps.n_max;
            stride = ps.stride;
            dFeat = ps.dFeat;
        } else {
        pass_tables_valid = false;
        std::vector<float> feat((size_t)Vtot * Fw);
        std::vector<int32_t> inst_graph((size_t)Vtot);
        for (int g = 0; g < G; ++g) {
            const double *f = ccn_graph_tables_features(gt[g]);
            const int V = molecule[g]->nVertices;
            assert(ccn_graph_tables_feature_width(gt[g]) == Fw);
            for (int64_t i = 0; i < (int64_t)V * Fw; ++i) feat[(size_t)vbase[g] * Fw + i] = (float)f[i];
            for (int v = 0; v < V; ++v) inst_graph[(size_t)(vbase[g] + v)] = g;
        }
        if ((int)lv.size() < L) lv.resize(L);
        for (int l = 1; l <= L; ++l) {
            int nm = 1;
            for (int g = 0; g < G; ++g)
                for (int v = 0; v < molecule[g]->nVertices; ++v) nm = std::max(nm, ccn_graph_tables_vertex(gt[g], l, v, NULL, NULL, NULL, NULL));
            n_max[l] = nm;
            stride[l] = (int64_t)nm * nm * w[l];
            std::vector<int64_t> f_off((size_t)Vtot * nm, 0);
            std::vector<int32_t> m((size_t)Vtot * nm, 1), pos((size_t)Vtot * nm * nm, -1), nn((size_t)Vtot);
            std::vector<float> adj((size_t)Vtot * nm * nm, 0.f);
            for (int g = 0; g < G; ++g)
                for (int v = 0; v < molecule[g]->nVertices; ++v) {
                    const float *a;
                    const int32_t *src, *mm, *pp;
                    const int n = ccn_graph_tables_vertex(gt[g], l, v, &a, &src, &mm, &pp);
                    const size_t i = (size_t)(vbase[g] + v);
                    nn[i] = n;
                    for (int k = 0; k < n * n; ++k) adj[i * nm * nm + k] = a[k];  // compact [n, n]
                    for (int s = 0; s < n; ++s) {
                        f_off[i * nm + s] = (vbase[g] + src[s]) * stride[l - 1];
                        m[i * nm + s] = mm[s];
                        for (int r = 0; r < n; ++r) pos[(i * nm + s) * nm + r] = pp[s * n + r];
                    }
                }
            LevelDevice &d = lv[l - 1];
            d.f_off.upload(f_off);
            d.m.upload(m);
            d.pos.upload(pos);
            d.adj.upload(adj);
            d.n.upload(nn);
        }
        dFeat = d_feat.upload(feat);
        ps.dPtr = d_instptr.upload(vbase);
        ps.dIG = d_instgraph.upload(inst_graph);
        pass_tables_key = key;
        pass_tables_valid = true;
        }  // !reuse
        // ---- parameters up ----
        std::vector<float> Ht((size_t)Fw * C), tmp;
        for (int c = 0; c < C; ++c)
            for (int k = 0; k < Fw; ++k) Ht[(size_t)k * C + c] = (float)H->value[H->index(c, k)];
        float *dHt = d_Ht.upload(Ht);
        std::vector<float> zero(C, 0.f);
        float *dZero = d_zero.upload(zero);
        for (int l = 1; l <= L; ++l) {
            to_float(level[l]->K, tmp);
            if (K_TRANSPOSED) transpose_in_place(tmp, w[l], 18 * w[l - 1]);  // device layout is always [18 C_in, C_out]
            lv[l - 1].K.upload(tmp);
            to_float(level[l]->b, tmp);
            lv[l - 1].b.upload(tmp);
        }
        // ---- forward ----
        float *pre0 = d_pre0.floats((size_t)Vtot * C), *act0 = d_act0.floats((size_t)Vtot * C);
        CCN_B200_CHECK(ctx, ccn_mix_forward(ctx, dFeat, dHt, dZero, pre0, act0, Vtot, Fw, C, alpha, NULL));  // level 0 (:563-573)
        const float *f_prev = act0;
        for (int l = 1; l <= L; ++l) {
            LevelDevice &d = lv[l - 1];
            const int nm = n_max[l];
            const size_t rows = (size_t)Vtot * nm * nm;
            const int Ci = w[l - 1], Co = w[l];
            float *X = d.X.floats(rows * 18 * Ci), *Y = d.Y.floats(rows * Co), *Z = d.Z.floats(rows * Co);
            // the contraction writes the n_i^2 real rows of every instance; the padding rows must read as zero in the mix.  A pass
            // over the same graphs as the last one finds them still zero (only the real rows are ever rewritten): no memset then
            if (!reuse) CCN_B200_CHECK(ctx, ccn_memset_zero(ctx, X, rows * 18 * Ci * sizeof(float), NULL));
            float *Tsc = NULL;
            if (!fuses(nm, Ci)) Tsc = d_T.floats((size_t)Vtot * nm * nm * nm * Ci);
            CCN_B200_CHECK(ctx, ccn_gather_level_forward(ctx, f_prev, static_cast<const int64_t *>(d.f_off.p), static_cast<const int32_t *>(d.m.p),
                                                         static_cast<const int32_t *>(d.pos.p), static_cast<const float *>(d.adj.p),
                                                         static_cast<const float *>(d.K.p), static_cast<const float *>(d.b.p), Tsc, X, Y, Z,
                                                         static_cast<const int32_t *>(d.n.p), nm, Ci, Co, Vtot, (int64_t)nm * nm,
                                                         CCN_ADJ_POSITIVE_PART, alpha, NULL));
            f_prev = Z;
        }
        ps.G = G;
        ps.Vtot = Vtot;
        ps.n_max = n_max;
        ps.stride = stride;
        ps.dFeat = dFeat, ps.dHt = dHt, ps.dZero = dZero, ps.pre0 = pre0, ps.act0 = act0, ps.f_last = f_prev;
        ps.shr0 = d_shr.floats((size_t)Vtot * w[0]);
        ps.dDgf = NULL;
        ps.g_cur = NULL;
    }

    // every level's feature into its columns of the concatenated [G, FW] graph feature (level 0: the [1,1,C] tensors); host copy out
    void level_features(std::vector<float> &h_feat) {
        ccn_ctx *ctx = context();
        const int L = nLevels, FW = multi_level_width(), G = ps.G;
        const std::vector<int> &w = width;
        float *gfeat = d_gfeat.floats((size_t)G * FW);
        int col = 0;
        for (int l = 0; l <= L; ++l) {
            const float *Zl = l == 0 ? ps.act0 : static_cast<const float *>(lv[l - 1].Z.p);
            const int32_t *nl = l == 0 ? NULL : static_cast<const int32_t *>(lv[l - 1].n.p);
            float *shl = l == 0 ? ps.shr0 : lv[l - 1].shr.floats((size_t)ps.Vtot * w[l]);
            CCN_B200_CHECK(ctx, ccn_level_features_forward(ctx, Zl, ps.stride[l], nl, ps.n_max[l], w[l], ps.Vtot, ps.dPtr, G, 0.01f, shl, gfeat + col, FW, NULL));
            col += w[l];
        }
        h_feat.resize((size_t)G * FW);
        CCN_B200_CHECK(ctx, ccn_d2h(ctx, &h_feat[0], gfeat, (size_t)G * FW * sizeof(float), NULL));
        CCN_B200_CHECK(ctx, ccn_stream_synchronize(ctx, NULL));
    }

    // dg: the gradient of the [G, FW] graph feature.  Writes the last level's share into the activation gradient of level L.
    void level_features_grad(const std::vector<float> &dg) {
        ccn_ctx *ctx = context();
        const int L = nLevels, FW = multi_level_width(), CL = width[L];
        ps.dDgf = d_dgf.upload(dg);
        ps.g_cur = L > 0 ? lv[L - 1].gZ.floats((size_t)ps.Vtot * ps.stride[L]) : d_gact0.floats((size_t)ps.Vtot * nChanels);
        const float *shL = L > 0 ? static_cast<const float *>(lv[L - 1].shr.p) : ps.shr0;
        const int32_t *nL = L > 0 ? static_cast<const int32_t *>(lv[L - 1].n.p) : NULL;
        CCN_B200_CHECK(ctx, ccn_level_features_backward(ctx, shL, ps.dDgf + (FW - CL), FW, ps.dIG, nL, ps.n_max[L], CL, ps.Vtot, 0.01f, ps.g_cur,
                                                        ps.stride[L], NULL));
    }

    // from ps.g_cur (the gradient of the level-L activations); with ps.dDgf set, every lower level's read-out share is folded in
    void levels_backward() {
        ccn_ctx *ctx = context();
        const int C = nChanels, L = nLevels, Fw = nFeatures * (nDepth + 1), FW = multi_level_width();
        const std::vector<int> &w = width;
        const float alpha = 0.01f;
        const int64_t Vtot = ps.Vtot;
        float *g_cur = ps.g_cur;
        for (int l = L; l >= 1; --l) {
            LevelDevice &d = lv[l - 1];
            const int nm = ps.n_max[l];
            const size_t rows = (size_t)Vtot * nm * nm;
            const int Ci = w[l - 1], Co = w[l];
            float *gK = d.gK.floats((size_t)18 * Ci * Co), *gb = d.gb.floats(Co);
            CCN_B200_CHECK(ctx, ccn_memset_zero(ctx, gK, (size_t)18 * Ci * Co * sizeof(float), NULL));
            CCN_B200_CHECK(ctx, ccn_memset_zero(ctx, gb, Co * sizeof(float), NULL));
            float *g_prev = l > 1 ? lv[l - 2].gZ.floats((size_t)Vtot * ps.stride[l - 1]) : d_gact0.floats((size_t)Vtot * C);
            if (ps.dDgf) {  // level l-1 feeds the read-out directly: its share initialises the gradient (padding rows = 0)
                int col = 0;
                for (int k = 0; k < l - 1; ++k) col += w[k];
                const float *shp = l > 1 ? static_cast<const float *>(lv[l - 2].shr.p) : ps.shr0;
                const int32_t *np = l > 1 ? static_cast<const int32_t *>(lv[l - 2].n.p) : NULL;
                CCN_B200_CHECK(ctx, ccn_level_features_backward(ctx, shp, ps.dDgf + col, FW, ps.dIG, np, ps.n_max[l - 1], Ci, Vtot, alpha, g_prev,
                                                                ps.stride[l - 1], NULL));
            } else {
                CCN_B200_CHECK(ctx, ccn_memset_zero(ctx, g_prev, (size_t)Vtot * ps.stride[l - 1] * sizeof(float), NULL));
            }
            float *gX = d_gX.floats(rows * 18 * Ci);
            float *Tsc = NULL;
            if (!fuses(nm, Ci)) Tsc = d_T.floats((size_t)Vtot * nm * nm * nm * Ci);
            CCN_B200_CHECK(ctx, ccn_gather_level_backward(ctx, g_cur, static_cast<const float *>(d.X.p), static_cast<const float *>(d.Y.p),
                                                          static_cast<const float *>(d.K.p), static_cast<const float *>(d.b.p),
                                                          static_cast<const float *>(d.adj.p), static_cast<const int64_t *>(d.f_off.p),
                                                          static_cast<const int32_t *>(d.m.p), static_cast<const int32_t *>(d.pos.p), gX, Tsc,
                                                          g_prev, gK, gb, static_cast<const int32_t *>(d.n.p), nm, Ci, Co, Vtot,
                                                          (int64_t)nm * nm, CCN_ADJ_POSITIVE_PART, alpha, NULL));
            add_gradient(level[l]->K, gK, K_TRANSPOSED ? 18 * Ci : 0, Co);
            add_gradient(level[l]->b, gb);
            g_cur = g_prev;
        }
        float *gHt = d_gHt.floats((size_t)Fw * C), *gdummy = d_zero2.floats(C);
        CCN_B200_CHECK(ctx, ccn_memset_zero(ctx, gHt, (size_t)Fw * C * sizeof(float), NULL));
        CCN_B200_CHECK(ctx, ccn_memset_zero(ctx, gdummy, C * sizeof(float), NULL));
        CCN_B200_CHECK(ctx, ccn_mix_backward(ctx, ps.dFeat, ps.dHt, ps.dZero, ps.pre0, g_cur, NULL, gHt, gdummy, Vtot, Fw, C, alpha, 0.f, NULL));
        std::vector<float> h((size_t)Fw * C);
        CCN_B200_CHECK(ctx, ccn_d2h(ctx, &h[0], gHt, h.size() * sizeof(float), NULL));
        CCN_B200_CHECK(ctx, ccn_stream_synchronize(ctx, NULL));
        for (int c = 0; c < C; ++c)
            for (int k = 0; k < Fw; ++k) H->gradient[H->index(c, k)] += h[(size_t)k * C + c];
    }

    // the sum of all level widths (the multi-level read-out's feature width)
    int multi_level_width() const {
        int t = 0;
        for (size_t l = 0; l < width.size(); ++l) t += width[l];
        return t;
    }

private:
    template <class TargetT>
    double pass(int G, DenseGraph **molecule, TargetT *target, bool need_grads, int out0) {
        ccn_ctx *ctx = context();
        const int C = nChanels, L = nLevels;
        const float alpha = 0.01f;
        levels_forward(G, molecule);
        const int FW = feature_width(), CL = width[L];
        std::vector<float> h_pred(G), h_loss(G), h_feat;
        std::vector<double> hid;  // hidden-layer head (host): pre-activations [G, nHidden]
        float *dW = NULL, *dT = NULL, *gfeat = NULL, *pred = NULL;
        const int32_t *nL = L > 0 ? static_cast<const int32_t *>(lv[L - 1].n.p) : NULL;
        if (head == HEAD_INNER_PRODUCT) {
            // ---- read-out + loss on the device (:620-639) ----
            std::vector<float> tmp, tgt(G, 0.f);
            to_float(W, tmp);
            dW = d_W.upload(tmp);
            if (target)
                for (int g = 0; g < G; ++g) tgt[g] = (float)target[g];
            dT = d_target.upload(tgt);
            gfeat = d_gfeat.floats((size_t)G * FW);
            pred = d_pred.floats(G);
            float *loss = d_loss.floats(G);
            CCN_B200_CHECK(ctx, ccn_readout_forward(ctx, ps.f_last, ps.stride[L], nL, ps.n_max[L], CL, ps.Vtot, ps.dPtr, G, dW, dT, alpha, ps.shr0, gfeat,
                                                    pred, loss, NULL));
            h_feat.resize((size_t)G * FW);
            CCN_B200_CHECK(ctx, ccn_d2h(ctx, &h_pred[0], pred, G * sizeof(float), NULL));
            CCN_B200_CHECK(ctx, ccn_d2h(ctx, &h_loss[0], loss, G * sizeof(float), NULL));
            CCN_B200_CHECK(ctx, ccn_d2h(ctx, &h_feat[0], gfeat, (size_t)G * FW * sizeof(float), NULL));
            CCN_B200_CHECK(ctx, ccn_stream_synchronize(ctx, NULL));
        } else {
            level_features(h_feat);
            // hidden = W1 gf (MatVecMul), LeakyReLU, predict = <., W2> (InnerProduct), SquaredLoss -- [G, FW] x [FW/2, FW]: host work
            const int nh = W2->size;
            hid.assign((size_t)G * nh, 0.0);
            for (int g = 0; g < G; ++g) {
                double p = 0.0;
                for (int i = 0; i < nh; ++i) {
                    double a = 0.0;
                    for (int k = 0; k < FW; ++k) a += W1->value[W1->index(i, k)] * (double)h_feat[(size_t)g * FW + k];
                    hid[(size_t)g * nh + i] = a;
                    p += (a > 0.0 ? a : (double)alpha * a) * W2->value[i];
                }
                h_pred[g] = (float)p;
                const double d = p - (target ? (double)target[g] : 0.0);
                h_loss[g] = (float)(0.5 * d * d);
            }
        }
        double total = 0.0;
        for (int g = 0; g < G; ++g) {
            last_predict[out0 + g] = h_pred[g];
            last_loss[out0 + g] = target ? h_loss[g] : 0.0;
            total += last_loss[out0 + g];
            for (int c = 0; c < FW; ++c) last_feature[(size_t)(out0 + g) * FW + c] = h_feat[(size_t)g * FW + c];
        }
        if (!need_grads) return total;
        // ---- backward ----
        if (head == HEAD_INNER_PRODUCT) {
            float *gW = d_gW.floats(C);
            CCN_B200_CHECK(ctx, ccn_memset_zero(ctx, gW, C * sizeof(float), NULL));
            ps.g_cur = L > 0 ? lv[L - 1].gZ.floats((size_t)ps.Vtot * ps.stride[L]) : d_gact0.floats((size_t)ps.Vtot * C);
            CCN_B200_CHECK(ctx, ccn_readout_backward(ctx, ps.shr0, gfeat, pred, dT, dW, ps.dIG, nL, ps.n_max[L], CL, ps.Vtot, G, alpha, ps.g_cur,
                                                     ps.stride[L], gW, NULL));
            levels_backward();
            add_gradient(W, gW);
        } else {
            const int nh = W2->size;
            std::vector<float> dg((size_t)G * FW, 0.f);
            for (int g = 0; g < G; ++g) {
                const double d = (double)h_pred[g] - (double)target[g];  // SquaredLoss backward
                for (int i = 0; i < nh; ++i) {
                    const double a = hid[(size_t)g * nh + i];
                    W2->gradient[i] += d * (a > 0.0 ? a : (double)alpha * a);               // InnerProduct
                    const double dh = d * W2->value[i] * (a > 0.0 ? 1.0 : (double)alpha);   // LeakyReLU
                    for (int k = 0; k < FW; ++k) {                                        // MatVecMul
                        W1->gradient[W1->index(i, k)] += dh * (double)h_feat[(size_t)g * FW + k];
                        dg[(size_t)g * FW + k] += (float)(dh * W1->value[W1->index(i, k)]);
                    }
                }
            }
            level_features_grad(dg);
            levels_backward();
        }
        return total;
    }

    static bool fuses(int n_max, int C) { return n_max <= 32 && (C == 8 || C == 16 || C == 32 || C == 64 || C == 128); }

    static void to_float(const Vector *v, std::vector<float> &out) {
        out.resize(v->size);
        for (int i = 0; i < v->size; ++i) out[i] = (float)v->value[i];
    }

    // rows x cols (row-major) -> cols x rows
    static void transpose_in_place(std::vector<float> &v, int rows, int cols) {
        std::vector<float> t(v.size());
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) t[(size_t)c * rows + r] = v[(size_t)r * cols + c];
        v.swap(t);
    }

    // param->gradient += the device array; dev_rows > 0: the device array is [dev_rows, dev_cols] and the parameter its transpose
    void add_gradient(Vector *param, const float *dev, int dev_rows = 0, int dev_cols = 0) {
        ccn_ctx *ctx = context();
        std::vector<float> h(param->size);
        CCN_B200_CHECK(ctx, ccn_d2h(ctx, &h[0], dev, h.size() * sizeof(float), NULL));
        CCN_B200_CHECK(ctx, ccn_stream_synchronize(ctx, NULL));
        if (dev_rows > 0) transpose_in_place(h, dev_rows, dev_cols);
        for (int i = 0; i < param->size; ++i) param->gradient[i] += h[i];
    }

    std::map<DenseGraph *, Cached> cache;
    std::vector<LevelDevice> lv;
    RawDevice d_feat, d_Ht, d_zero, d_zero2, d_pre0, d_act0, d_gact0, d_gHt, d_W, d_gW, d_target, d_shr, d_gfeat, d_pred, d_loss, d_instptr,
        d_instgraph, d_gX, d_T, d_dgf;
};

// SMP_beta (GraphFlow/SMP_beta.h): Adam, K_l stored [18 C, C].
class SMP_beta : public SMP_model<Adam, false> {
public:
    SMP_beta(int max_nVertices, int nLevels, int nChanels, int nFeatures, int nDepth)
        : SMP_model<Adam, false>(max_nVertices, nLevels, nChanels, nFeatures, nDepth, new Adam()) {}
};

// SMP_2D_ver8 (GraphFlow/SMP_2D_ver8.h:32-43; BASELINE config 4's model): the same wiring with the feature mix done by
// CustomMatMulTensor (K_l stored [C, 18 C], :130, 526-527) and the Momentum optimizer (:205).
class SMP_2D_ver8 : public SMP_model<Momentum, true> {
public:
    SMP_2D_ver8(int max_nVertices, int nLevels, int nChanels, int nFeatures, int nDepth, double momentum_param)
        : SMP_model<Momentum, true>(max_nVertices, nLevels, nChanels, nFeatures, nDepth, new Momentum(momentum_param)) {}
};

// SMP_omega (GraphFlow/SMP_omega.h:29-1247): SMP_beta's wiring with receptive fields limited to max_receptive_field members
// (limit_receptive_field, :476-510: nearest first, ties by WL rank, whole outermost shells dropped; then ordered by rank).
// Same calls as SMP_beta; only the graph tables differ (CCN_GRAPH_OMEGA_WL).
class SMP_omega : public SMP_beta {
public:
    SMP_omega(int max_nVertices, int max_receptive_field, int nLevels, int nChanels, int nFeatures, int nDepth)
        : SMP_beta(max_nVertices, nLevels, nChanels, nFeatures, nDepth) {
        assert(max_receptive_field <= max_nVertices);
        this->graph_kind = CCN_GRAPH_OMEGA_WL;
        this->max_receptive_field = max_receptive_field;
    }
};

// SMP_omega_physics (GraphFlow/SMP_omega_physics.h; BASELINE config 3's model): see the PhysicsTag constructor above.
class SMP_omega_physics : public SMP_model<Adam, false> {
public:
    SMP_omega_physics(int max_nVertices, int max_receptive_field, int nLevels, int nChanels, int nFeatures)
        : SMP_model<Adam, false>(PhysicsTag(), max_nVertices, max_receptive_field, nLevels, nChanels, nFeatures, new Adam()) {
        assert(max_receptive_field <= max_nVertices);
    }
};

// SMP_omega_pairgraphs (GraphFlow/SMP_omega_pairgraphs.h:29-1274): one example is a pair (graph, line graph); the SMP_omega_physics
// path runs once on each with separate parameters (`computation_graph_`, :147-281), the level features are concatenated level by
// level (:705-710) and two hidden layers follow (MatVecMul + LeakyReLU twice, InnerProduct, SquaredLoss, :714-729).  Parameters in
// the reference's registration order (:365-377): H_1, H_2, (K1_l, b1_l, K2_l, b2_l)..., W1, W2, W3, initialised from rand() in that
// order like the reference (same seed => same parameters).  The two trunks run on the device; the head (a [graphs, Ctot] matrix
// against [Ctot/2, Ctot] and [Ctot/4, Ctot/2] weights) on the host.
class SMP_omega_pairgraphs {
public:
    typedef SMP_model<Adam, false> Trunk;

    SMP_omega_pairgraphs(int max_nVertices_1, int max_nVertices_2, int max_receptive_field, int nLevels, int nChanels, int nFeatures_1,
                         int nFeatures_2) {
        assert(max_receptive_field <= max_nVertices_1 && max_receptive_field <= max_nVertices_2);
        this->nLevels = nLevels;
        chunk_graphs = 256;
        sgd = new Adam();
        t1 = new Trunk(Trunk::PhysicsTag(), max_nVertices_1, max_receptive_field, nLevels, nChanels, nFeatures_1, sgd, false);
        t2 = new Trunk(Trunk::PhysicsTag(), max_nVertices_2, max_receptive_field, nLevels, nChanels, nFeatures_2, sgd, false);
        const int total = t1->multi_level_width() + t2->multi_level_width();
        const int h1 = std::max(total / 2, 10), h2 = std::max(h1 / 2, 10);  // :328-329
        W1 = new Matrix(h1, total);
        W2 = new Matrix(h2, h1);
        W3 = new Vector(h2);
        sgd->add(t1->H);
        sgd->add(t2->H);
        for (int l = 1; l <= nLevels; ++l) {
            sgd->add(t1->level[l]->K);
            sgd->add(t1->level[l]->b);
            sgd->add(t2->level[l]->K);
            sgd->add(t2->level[l]->b);
        }
        sgd->add(W1);
        sgd->add(W2);
        sgd->add(W3);
        for (size_t i = 0; i < sgd->params.size(); ++i) {  // GraphFlow::uniform_init (GraphFlow.h:1297-1306), registration order
            Vector *V = sgd->params[i];
            for (int j = 0; j < V->size; ++j) {
                V->value[j] = (double)(rand() % 10) / (10.0 * V->size);
                if (rand() % 2 == 1) V->value[j] = -V->value[j];
            }
        }
    }

    std::pair<double, double> BatchLearn(int nBatch, DenseGraph **molecule_1, DenseGraph **molecule_2, double *target, double learning_rate) {
        std::pair<double, double> ret;
        ret.first = run(nBatch, molecule_1, molecule_2, target, true);
        sgd->Learn(learning_rate, nBatch);
        ret.second = run(nBatch, molecule_1, molecule_2, target, false);
        return ret;
    }
    double getLoss(int nBatch, DenseGraph **molecule_1, DenseGraph **molecule_2, double *target) {
        return run(nBatch, molecule_1, molecule_2, target, false);
    }
    double Predict(DenseGraph *molecule_1, DenseGraph *molecule_2) {
        run(1, &molecule_1, &molecule_2, NULL, false);
        return last_predict[0];
    }
    void save_model(std::string filename) {
        std::ofstream file(filename.c_str(), std::ios::out);
        for (size_t i = 0; i < sgd->params.size(); ++i)
            for (int j = 0; j < sgd->params[i]->size; ++j) file << sgd->params[i]->value[j] << " ";
        file.close();
    }
    void load_model(std::string filename) {
        std::ifstream file(filename.c_str(), std::ios::in);
        for (size_t i = 0; i < sgd->params.size(); ++i)
            for (int j = 0; j < sgd->params[i]->size; ++j) file >> sgd->params[i]->value[j];
        file.close();
    }
    void release() {
        t1->release();
        t2->release();
    }

    int nLevels, chunk_graphs;
    Trunk *t1, *t2;
    Matrix *W1, *W2;
    Vector *W3;
    Adam *sgd;
    std::vector<double> last_predict, last_loss;

private:
    double run(int nBatch, DenseGraph **m1, DenseGraph **m2, double *target, bool need_grads) {
        last_predict.assign(nBatch, 0.0);
        last_loss.assign(nBatch, 0.0);
        if (need_grads)
            for (size_t i = 0; i < sgd->params.size(); ++i)
                for (int j = 0; j < sgd->params[i]->size; ++j) sgd->params[i]->gradient[j] = 0.0;
        const double alpha = 0.01;
        const int L = nLevels, F1 = t1->multi_level_width(), F2 = t2->multi_level_width(), FW = F1 + F2;
        const int h1n = W2->nColumns, h2n = W3->size;
        double total = 0.0;
        for (int g0 = 0; g0 < nBatch; g0 += chunk_graphs) {
            const int G = std::min(chunk_graphs, nBatch - g0);
            std::vector<float> f1, f2;
            t1->levels_forward(G, m1 + g0);
            t1->level_features(f1);
            t2->levels_forward(G, m2 + g0);
            t2->level_features(f2);
            std::vector<float> dg1((size_t)G * F1, 0.f), dg2((size_t)G * F2, 0.f);
            std::vector<double> gf(FW), a1(h1n), z1(h1n), a2(h2n), z2(h2n), dz1(h1n), dz2(h2n), dgf(FW);
            for (int g = 0; g < G; ++g) {
                // ConcatVectors level by level: [lf1_0, lf2_0, lf1_1, lf2_1, ...]
                int c1 = 0, c2 = 0, c = 0;
                for (int l = 0; l <= L; ++l) {
                    for (int k = 0; k < t1->width[l]; ++k) gf[c++] = f1[(size_t)g * F1 + c1 + k];
                    for (int k = 0; k < t2->width[l]; ++k) gf[c++] = f2[(size_t)g * F2 + c2 + k];
                    c1 += t1->width[l];
                    c2 += t2->width[l];
                }
                for (int i = 0; i < h1n; ++i) {
                    double a = 0.0;
                    for (int k = 0; k < FW; ++k) a += W1->value[W1->index(i, k)] * gf[k];
                    z1[i] = a;
                    a1[i] = a > 0.0 ? a : alpha * a;
                }
                double p = 0.0;
                for (int i = 0; i < h2n; ++i) {
                    double a = 0.0;
                    for (int k = 0; k < h1n; ++k) a += W2->value[W2->index(i, k)] * a1[k];
                    z2[i] = a;
                    a2[i] = a > 0.0 ? a : alpha * a;
                    p += a2[i] * W3->value[i];
                }
                last_predict[g0 + g] = p;
                if (!target) continue;
                const double d = p - target[g0 + g];
                last_loss[g0 + g] = 0.5 * d * d;
                total += last_loss[g0 + g];
                if (!need_grads) continue;
                for (int k = 0; k < h1n; ++k) dz1[k] = 0.0;
                for (int i = 0; i < h2n; ++i) {
                    W3->gradient[i] += d * a2[i];
                    dz2[i] = d * W3->value[i] * (z2[i] > 0.0 ? 1.0 : alpha);
                    for (int k = 0; k < h1n; ++k) {
                        W2->gradient[W2->index(i, k)] += dz2[i] * a1[k];
                        dz1[k] += dz2[i] * W2->value[W2->index(i, k)];
                    }
                }
                for (int k = 0; k < FW; ++k) dgf[k] = 0.0;
                for (int i = 0; i < h1n; ++i) {
                    const double dh = dz1[i] * (z1[i] > 0.0 ? 1.0 : alpha);
                    for (int k = 0; k < FW; ++k) {
                        W1->gradient[W1->index(i, k)] += dh * gf[k];
                        dgf[k] += dh * W1->value[W1->index(i, k)];
                    }
                }
                c1 = c2 = c = 0;
                for (int l = 0; l <= L; ++l) {
                    for (int k = 0; k < t1->width[l]; ++k) dg1[(size_t)g * F1 + c1 + k] = (float)dgf[c++];
                    for (int k = 0; k < t2->width[l]; ++k) dg2[(size_t)g * F2 + c2 + k] = (float)dgf[c++];
                    c1 += t1->width[l];
                    c2 += t2->width[l];
                }
            }
            if (need_grads) {
                t1->level_features_grad(dg1);
                t1->levels_backward();
                t2->level_features_grad(dg2);
                t2->levels_backward();
            }
        }
        return total;
    }
};

}  // namespace ccn_b200

#ifdef CCN_B200_DROP_IN
typedef ccn_b200::SMP_omega_pairgraphs SMP_omega_pairgraphs;
typedef ccn_b200::SMP_omega_physics SMP_omega_physics;
typedef ccn_b200::SMP_beta SMP_beta;
typedef ccn_b200::SMP_omega SMP_omega;
typedef ccn_b200::SMP_2D_ver8 SMP_2D_ver8;
#endif

#endif  // GRAPHFLOW_B200_SMP_BETA_B200_H_INCLUDED
