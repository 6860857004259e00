// oracle/ref_model_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// extern "C" shim around the UNMODIFIED reference models GraphFlow/SMP_beta.h and GraphFlow/SMP_2D_ver8.h (second-order
// CCN, BASELINE.json configs 2 and 4), compiled with -I/root/reference/GraphFlow by oracle/Makefile into oracle/_ref/libgfref_model_f64.so.  It drives
// the model exactly as SMP_beta::BatchLearn does for one example (SMP_beta.h:757-765: complete_computation_graph,
// target, graph->forward(), graph->backward()) with caller-supplied parameters, and exports what a drop-in
// implementation must reproduce: the graph feature (SMP_beta::Feature, :931-943), the loss, every parameter gradient,
// and the receptive fields phi_l(v) the model derived from the graph (:461-489).
#include <cstdlib>
#include <iostream>
#include <cstring>
#include <algorithm>
#include <vector>

#include "SMP_beta.h"
#include "SMP_2D_ver8.h"
#include "SMP_omega_physics.h"
#include "SMP_omega.h"
#include "SMP_omega_pairgraphs.h"
#include "Momentum.h"

namespace {

template <class Model>
int run_model(Model *model, int V, const int *adj, const double *feat, int L, int C, int F, const double *params, double target,
              double *graph_feature, double *loss, double *grads, int *phi_out) {
    DenseGraph *g = new DenseGraph(V, F);
    for (int i = 0; i < V; ++i) {
        for (int j = 0; j < V; ++j) g->adj[i][j] = adj[i * V + j];
        for (int f = 0; f < F; ++f) g->feature[i][f] = feat[i * F + f];
    }
    int total = 0;
    for (size_t p = 0; p < model->sgd->params.size(); ++p) {
        Vector *v = model->sgd->params[p];
        if (params) std::memcpy(v->value, params + total, sizeof(double) * v->size);
        total += v->size;
    }
    model->complete_computation_graph(g);
    model->target->value[0] = target;
    model->graph->forward();
    model->graph->backward();
    for (int c = 0; c < C; ++c) graph_feature[c] = model->graph_feature->value[c];
    *loss = model->sql->getLoss();
    int off = 0;
    for (size_t p = 0; p < model->sgd->params.size(); ++p) {
        Vector *v = model->sgd->params[p];
        std::memcpy(grads + off, v->gradient, sizeof(double) * v->size);
        off += v->size;
    }
    if (phi_out) {
        for (int l = 0; l <= L; ++l)
            for (int v = 0; v < V; ++v) {
                int *dst = phi_out + ((size_t)l * V + v) * (V + 1);
                dst[0] = (int)model->level[l]->phi[v].size();
                for (int i = 0; i < dst[0]; ++i) dst[1 + i] = model->level[l]->phi[v][i];
            }
    }
    return total;  // the model and the graph are leaked on purpose: the reference has no usable destructor
}

}  // namespace

extern "C" {

// params / grads: flat, in the optimizer's registration order (SMP_beta.h:276-282): H [C, F (nDepth+1)], then per
// level l = 1..L: K_l [18 C, C], b_l [C]; then W [C].
// phi_out: [(L+1)][V][V+1] ints: count followed by the members of phi_l(v) in the model's order.
// returns the number of parameter scalars.
int gfref_smp_beta_f64(int V, const int *adj, const double *feat, int L, int C, int F, int nDepth, const double *params,
                       double target, double *graph_feature, double *loss, double *grads, int *phi_out) {
    srand(1);
    // max_nVertices >= nFeatures: the reference sizes its WL histogram rows by max_nVertices * (nDepth + 1) but fills
    // nFeatures * (nDepth + 1) entries (SMP_beta.h:307-309 vs :367-389) -- a heap overflow for graphs with fewer vertices
    // than features (found on the 3-atom H2O molecule, whose loss then depended on the heap layout).  Its own tests use
    // max_nVertices = 10 >= nFeatures = 4 (tests/test_SMP_beta.cpp:22-26), the regime reproduced here.
    return run_model(new SMP_beta(std::max(V, F), L, C, F, nDepth), V, adj, feat, L, C, F, params, target, graph_feature, loss, grads, phi_out);
}

// SMP_2D_ver8 (SMP_2D_ver8.h; BASELINE.json config 4's model): the same wiring with the mix done by CustomMatMulTensor,
// i.e. K_l stored [C, 18 C] (SMP_2D_ver8.h:130, 526-527); parameter order H, K_l, b_l, W (:205-211).
int gfref_smp_2d_ver8_f64(int V, const int *adj, const double *feat, int L, int C, int F, int nDepth, const double *params,
                          double target, double *graph_feature, double *loss, double *grads, int *phi_out) {
    srand(1);
    return run_model(new SMP_2D_ver8(std::max(V, F), L, C, F, nDepth, 0.9), V, adj, feat, L, C, F, params, target, graph_feature, loss, grads,
                     phi_out);
}

// SMP_omega_physics (SMP_omega_physics.h; BASELINE.json config 3's model): channels halve per level (:142-146), receptive
// fields limited to max_field members (:367-418), every level feeds the read-out, which ends in a hidden layer (:560-595).
// Parameter order (:256-262): H [C, F], K_l [18 C_{l-1}, C_l], b_l [C_l] for l = 1..L, W1 [Ctot/2, Ctot], W2 [Ctot/2].
// graph_feature has Ctot = sum of the level widths entries (pass C = Ctot for the output size).
int gfref_smp_omega_physics_f64(int V, const int *adj, const double *feat, int max_field, int L, int C, int F, int Ctot,
                                const double *params, double target, double *graph_feature, double *loss, double *grads,
                                int *phi_out) {
    srand(1);
    return run_model(new SMP_omega_physics(V, max_field, L, C, F), V, adj, feat, L, Ctot, F, params, target, graph_feature, loss,
                     grads, phi_out);
}

// SMP_omega (SMP_omega.h): SMP_beta's wiring with receptive fields limited to max_field members (:476-531).
int gfref_smp_omega_f64(int V, const int *adj, const double *feat, int max_field, int L, int C, int F, int nDepth, const double *params,
                        double target, double *graph_feature, double *loss, double *grads, int *phi_out) {
    srand(1);
    return run_model(new SMP_omega(std::max(V, F), max_field, L, C, F, nDepth), V, adj, feat, L, C, F, params, target, graph_feature,
                     loss, grads, phi_out);
}

// SMP_omega_pairgraphs (SMP_omega_pairgraphs.h): one example = (graph, line graph); the path runs once on each with separate
// parameters, the level features are concatenated level by level and a two-hidden-layer head follows (:657-730).
// params / grads in registration order (:365-377): H_1, H_2, per level K1_l, b1_l, K2_l, b2_l, then W1, W2, W3.
// graph_feature: Ctot = 2 * sum of the level widths entries.  Returns the number of parameter scalars.
int gfref_smp_omega_pairgraphs_f64(int V1, const int *adj1, const double *feat1, int F1, int V2, const int *adj2, const double *feat2,
                                   int F2, int max_field, int L, int C, int Ctot, const double *params, double target,
                                   double *graph_feature, double *loss, double *predict, double *grads) {
    srand(1);
    SMP_omega_pairgraphs *model = new SMP_omega_pairgraphs(V1, V2, max_field, L, C, F1, F2);
    DenseGraph *g1 = new DenseGraph(V1, F1), *g2 = new DenseGraph(V2, F2);
    for (int i = 0; i < V1; ++i) {
        for (int j = 0; j < V1; ++j) g1->adj[i][j] = adj1[i * V1 + j];
        for (int f = 0; f < F1; ++f) g1->feature[i][f] = feat1[i * F1 + f];
    }
    for (int i = 0; i < V2; ++i) {
        for (int j = 0; j < V2; ++j) g2->adj[i][j] = adj2[i * V2 + j];
        for (int f = 0; f < F2; ++f) g2->feature[i][f] = feat2[i * F2 + f];
    }
    int total = 0;
    for (size_t p = 0; p < model->sgd->params.size(); ++p) {
        Vector *v = model->sgd->params[p];
        if (params) std::memcpy(v->value, params + total, sizeof(double) * v->size);
        total += v->size;
    }
    model->complete_computation_graph(g1, g2);
    model->target->value[0] = target;
    model->graph->forward();
    model->graph->backward();
    for (int c = 0; c < Ctot; ++c) graph_feature[c] = model->graph_feature->value[c];
    *loss = model->sql->getLoss();
    *predict = model->predict->value[0];
    int off = 0;
    for (size_t p = 0; p < model->sgd->params.size(); ++p) {
        Vector *v = model->sgd->params[p];
        std::memcpy(grads + off, v->gradient, sizeof(double) * v->size);
        off += v->size;
    }
    return total;
}

// SMP_beta::save_model / load_model (SMP_beta.h:980-1002).  `params` are written into the model and saved to `save_path`
// (if non-NULL); then `load_path` (if non-NULL) is loaded and the model's parameters are returned in `loaded`.
int gfref_smp_beta_checkpoint_f64(int Vmax, int L, int C, int F, int nDepth, const double *params, const char *save_path,
                                  const char *load_path, double *loaded) {
    SMP_beta *model = new SMP_beta(std::max(Vmax, F), L, C, F, nDepth);
    int total = 0;
    for (size_t p = 0; p < model->sgd->params.size(); ++p) {
        Vector *v = model->sgd->params[p];
        if (params) std::memcpy(v->value, params + total, sizeof(double) * v->size);
        total += v->size;
    }
    if (save_path) model->save_model(save_path);
    if (load_path) {
        model->load_model(load_path);
        int off = 0;
        for (size_t p = 0; p < model->sgd->params.size(); ++p) {
            Vector *v = model->sgd->params[p];
            std::memcpy(loaded + off, v->value, sizeof(double) * v->size);
            off += v->size;
        }
    }
    return total;
}

// SMP_beta::BatchLearn (SMP_beta.h:745-772): `epochs` calls on the same batch of graphs.  adj / feat hold the graphs back
// to back ([V_i, V_i] ints, [V_i, F] doubles).  losses[2e], losses[2e+1] = the summed loss before / after epoch e
// (BatchLearn's return value); params_out = the parameters after the last epoch.
int gfref_smp_beta_batchlearn_f64(int nGraphs, const int *V, const int *adj, const double *feat, int L, int C, int F, int nDepth,
                                  const double *params, const double *targets, int epochs, double lr, double *losses,
                                  double *params_out) {
    int Vmax = F;
    for (int i = 0; i < nGraphs; ++i) Vmax = std::max(Vmax, V[i]);
    SMP_beta *model = new SMP_beta(Vmax, L, C, F, nDepth);
    DenseGraph **graphs = new DenseGraph *[nGraphs];
    size_t ao = 0, fo = 0;
    for (int gi = 0; gi < nGraphs; ++gi) {
        DenseGraph *g = new DenseGraph(V[gi], F);
        for (int i = 0; i < V[gi]; ++i) {
            for (int j = 0; j < V[gi]; ++j) g->adj[i][j] = adj[ao + (size_t)i * V[gi] + j];
            for (int f = 0; f < F; ++f) g->feature[i][f] = feat[fo + (size_t)i * F + f];
        }
        ao += (size_t)V[gi] * V[gi];
        fo += (size_t)V[gi] * F;
        graphs[gi] = g;
    }
    int total = 0;
    for (size_t p = 0; p < model->sgd->params.size(); ++p) {
        Vector *v = model->sgd->params[p];
        std::memcpy(v->value, params + total, sizeof(double) * v->size);
        total += v->size;
    }
    std::vector<double> tg(targets, targets + nGraphs);
    for (int e = 0; e < epochs; ++e) {
        std::pair<double, double> r = model->BatchLearn(nGraphs, graphs, &tg[0], lr);
        losses[2 * e] = r.first;
        losses[2 * e + 1] = r.second;
    }
    int off = 0;
    for (size_t p = 0; p < model->sgd->params.size(); ++p) {
        Vector *v = model->sgd->params[p];
        std::memcpy(params_out + off, v->value, sizeof(double) * v->size);
        off += v->size;
    }
    return total;
}

// Adam (Adam.h) / Momentum (Momentum.h) on two registered parameter vectors of sizes n0 and n1 (values and per-step
// gradients packed back to back).  `steps` updates with grads[s * (n0+n1) ...]; mode 0: Adam::Learn(alpha, nBatch),
// 1: Adam::Learn(alpha), 2: Momentum::Learn(alpha, nBatch).
void gfref_optimizer_f64(int mode, double *values, const double *grads, int n0, int n1, int steps, double alpha, int nBatch) {
    Vector *a = new Vector(n0), *b = new Vector(n1);
    std::memcpy(a->value, values, sizeof(double) * n0);
    std::memcpy(b->value, values + n0, sizeof(double) * n1);
    Adam *adam = new Adam();
    Momentum *mom = new Momentum();
    adam->add(a);
    adam->add(b);
    mom->add(a);
    mom->add(b);
    for (int s = 0; s < steps; ++s) {
        const double *g = grads + (size_t)s * (n0 + n1);
        std::memcpy(a->gradient, g, sizeof(double) * n0);
        std::memcpy(b->gradient, g + n0, sizeof(double) * n1);
        if (mode == 0) adam->Learn(alpha, nBatch);
        else if (mode == 1) adam->Learn(alpha);
        else mom->Learn(alpha, nBatch);
    }
    std::memcpy(values, a->value, sizeof(double) * n0);
    std::memcpy(values + n0, b->value, sizeof(double) * n1);
}

}  // extern "C"
