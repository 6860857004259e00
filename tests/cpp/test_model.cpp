// tests/cpp/test_model.cpp -- the C++ model facade ccn_b200::SMP_beta (include/graphflow_b200/SMP_beta_b200.h) against the
// UNMODIFIED reference model ::SMP_beta (GraphFlow/SMP_beta.h, double tree), both constructed from the same srand() seed, on
// the four hard-coded molecules of the reference's tests/test_SMP_beta.cpp:70-146 (CH4, NH3, H2O, C2H4).  The recipe is the
// one of tests/test_SMP_similarity.cu:92-147 (same-seed models, compare Feature()) extended to Predict, getLoss and a few
// epochs of BatchLearn (same Adam object type on both sides).
//
//   test_model parity L C D      prints `model <what> err=... ok|FAIL` lines and `model failures=N`
//   test_model bench  L C V B    times BatchLearn's forward + backward on B random molecular graphs of V vertices: one JSON line
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "SMP_beta.h"   // the reference models (double tree)
#include "SMP_omega.h"
#include "SMP_2D_ver8.h"
#include "SMP_omega_physics.h"
#include "SMP_omega_pairgraphs.h"

#include "graphflow_b200/SMP_beta_b200.h"

static int failures = 0;
static void check(const char *what, double err, double tol) {
    const bool ok = err <= tol;
    std::printf("model %s err=%.3e tol=%.1e %s\n", what, err, tol, ok ? "ok" : "FAIL");
    if (!ok) ++failures;
}

static DenseGraph *molecule(const std::vector<std::string> &label, const std::vector<std::pair<int, int> > &edges) {
    DenseGraph *g = new DenseGraph((int)label.size(), 4);
    for (size_t i = 0; i < edges.size(); ++i) {
        g->adj[edges[i].first][edges[i].second] = 1;
        g->adj[edges[i].second][edges[i].first] = 1;
    }
    for (size_t v = 0; v < label.size(); ++v) {
        const std::string &s = label[v];
        g->feature[v][s == "C" ? 0 : s == "H" ? 1 : s == "N" ? 2 : 3] = 1.0;
    }
    return g;
}

static std::vector<DenseGraph *> four_molecules() {
    typedef std::pair<int, int> E;
    std::vector<DenseGraph *> out;
    out.push_back(molecule({"C", "H", "H", "H", "H"}, {E(0, 1), E(0, 2), E(0, 3), E(0, 4)}));                         // CH4
    out.push_back(molecule({"N", "H", "H", "H"}, {E(0, 1), E(0, 2), E(0, 3)}));                                        // NH3
    out.push_back(molecule({"O", "H", "H"}, {E(0, 1), E(0, 2)}));                                                      // H2O
    out.push_back(molecule({"C", "C", "H", "H", "H", "H"}, {E(0, 1), E(0, 2), E(0, 3), E(1, 4), E(1, 5)}));            // C2H4
    return out;
}

static double rel(double a, double b, double scale) { return std::fabs(a - b) / std::max(scale, 1e-12); }

static void parity(int L, int C, int D) {
    const int maxV = 10, F = 4, seed = 20171;
    srand(seed);
    SMP_beta *ref = new SMP_beta(maxV, L, C, F, D);
    srand(seed);
    ccn_b200::SMP_beta *mine = new ccn_b200::SMP_beta(maxV, L, C, F, D);
    double dp = 0;
    for (size_t i = 0; i < ref->sgd->params.size(); ++i)
        for (int j = 0; j < ref->sgd->params[i]->size; ++j)
            dp = std::max(dp, std::fabs(ref->sgd->params[i]->value[j] - mine->sgd->params[i]->value[j]));
    check("same_seed_parameters", dp, 0.0);

    std::vector<DenseGraph *> mol = four_molecules();
    double targets[4] = {-0.8, 0.3, 1.1, -0.2};
    double worst_f = 0, worst_p = 0;
    for (size_t i = 0; i < mol.size(); ++i) {
        std::vector<double> fr = ref->Feature(mol[i]), fm = mine->Feature(mol[i]);
        double sc = 0;
        for (size_t c = 0; c < fr.size(); ++c) sc = std::max(sc, std::fabs(fr[c]));
        for (size_t c = 0; c < fr.size(); ++c) worst_f = std::max(worst_f, rel(fr[c], fm[c], sc));
        const double pr = ref->Predict(mol[i]), pm = mine->Predict(mol[i]);
        worst_p = std::max(worst_p, rel(pr, pm, std::fabs(pr)));
    }
    check("Feature", worst_f, 1e-4);
    check("Predict", worst_p, 1e-4);
    const double lr0 = ref->getLoss(4, &mol[0], targets), lm0 = mine->getLoss(4, &mol[0], targets);
    check("getLoss", rel(lr0, lm0, lr0), 1e-4);

    // From here on both models get the same GENERIC parameters.  uniform_init draws multiples of 1/(10 size) (a tenth of them
    // exactly 0) and the atom features are one-hot / small counts, so many level-0 pre-activations H x_v are EXACTLY zero in
    // exact arithmetic -- the kink of the leaky ReLU, where the reference's double sum and any fp32 sum can land on different
    // sides (derivative 1 vs 0.01).  That is a property of the degenerate initial point, not of the path; one optimizer step
    // away from it the pre-activations are generic.
    for (size_t i = 0; i < ref->sgd->params.size(); ++i)
        for (int j = 0; j < ref->sgd->params[i]->size; ++j) {
            const double bump = 0.02 * (rand() / (RAND_MAX + 1.0) - 0.5) / std::sqrt((double)ref->sgd->params[i]->size);
            ref->sgd->params[i]->value[j] += bump;
            mine->sgd->params[i]->value[j] = ref->sgd->params[i]->value[j];
        }

    // gradients of one forward + backward, summed over the four molecules, parameter by parameter (H, K_1, b_1, ..., W)
    ref->sum_gradients->reset_sum_gradients();
    for (int i = 0; i < 4; ++i) {
        ref->complete_computation_graph(mol[i]);
        ref->target->value[0] = targets[i];
        ref->graph->forward();
        ref->graph->backward();
        ref->sum_gradients->cache_gradients();
    }
    ref->sum_gradients->get_sum_gradients();
    mine->Gradients(4, &mol[0], targets);
    for (size_t i = 0; i < ref->sgd->params.size(); ++i) {
        double d = 0, sc = 0;
        for (int j = 0; j < ref->sgd->params[i]->size; ++j) {
            d = std::max(d, std::fabs(ref->sgd->params[i]->gradient[j] - mine->sgd->params[i]->gradient[j]));
            sc = std::max(sc, std::fabs(ref->sgd->params[i]->gradient[j]));
        }
        if (i == 0 && d / std::max(sc, 1e-30) > 1e-4) {
            std::printf("model debug H gradient [c=0..1][k]: ref | mine\n");
            const int Fw = ref->sgd->params[0]->size / C;
            for (int c = 0; c < 2; ++c)
                for (int k = 0; k < Fw; ++k)
                    std::printf("   c=%d k=%d  % .6e  % .6e\n", c, k, (double)ref->sgd->params[0]->gradient[c * Fw + k],
                                (double)mine->sgd->params[0]->gradient[c * Fw + k]);
        }
        char name[64];
        std::snprintf(name, sizeof(name), "gradient_param_%d(size=%d)", (int)i, ref->sgd->params[i]->size);
        check(name, d / std::max(sc, 1e-30), 1e-4);
    }

    // a few epochs of BatchLearn: the same loss pair every epoch and the same parameters at the end
    double worst_l = 0;
    for (int e = 0; e < 5; ++e) {
        std::pair<double, double> a = ref->BatchLearn(4, &mol[0], targets, 0.001), b = mine->BatchLearn(4, &mol[0], targets, 0.001);
        worst_l = std::max(worst_l, std::max(rel(a.first, b.first, a.first), rel(a.second, b.second, a.second)));
    }
    check("BatchLearn_losses", worst_l, 2e-4);
    double dq = 0, sq = 0;
    for (size_t i = 0; i < ref->sgd->params.size(); ++i)
        for (int j = 0; j < ref->sgd->params[i]->size; ++j) {
            dq = std::max(dq, std::fabs(ref->sgd->params[i]->value[j] - mine->sgd->params[i]->value[j]));
            sq = std::max(sq, std::fabs(ref->sgd->params[i]->value[j]));
        }
    check("BatchLearn_parameters", dq / sq, 1e-3);

    // checkpoint: ours -> the reference's load_model -> same predictions
    mine->save_model("/tmp/ccn_b200_model.dat");
    srand(seed + 1);
    SMP_beta *ref2 = new SMP_beta(maxV, L, C, F, D);
    ref2->load_model("/tmp/ccn_b200_model.dat");
    double worst_c = 0;
    for (size_t i = 0; i < mol.size(); ++i) {
        const double pr = ref2->Predict(mol[i]), pm = mine->Predict(mol[i]);
        worst_c = std::max(worst_c, rel(pr, pm, std::fabs(pr)));
    }
    check("checkpoint_roundtrip", worst_c, 1e-4);
    std::printf("model kernel_launches=%lld\n", mine->kernel_launches());
    mine->release();
}

static DenseGraph *random_molecular_graph(int V);

// ::SMP_omega vs ccn_b200::SMP_omega on graphs LARGER than the receptive-field limit (so limit_receptive_field cuts), same
// seed, then generic parameters: Predict / Feature / getLoss / every gradient / three epochs of BatchLearn.
static void parity_omega(int L, int C, int D, int max_field) {
    const int maxV = 12, F = 4, seed = 777;
    srand(seed);
    SMP_omega *ref = new SMP_omega(maxV, max_field, L, C, F, D);
    srand(seed);
    ccn_b200::SMP_omega *mine = new ccn_b200::SMP_omega(maxV, max_field, L, C, F, D);
    std::vector<DenseGraph *> mol;
    srand(5);
    for (int i = 0; i < 4; ++i) mol.push_back(random_molecular_graph(8 + i));
    double targets[4] = {0.4, -0.7, 1.2, 0.1};
    for (size_t i = 0; i < ref->sgd->params.size(); ++i)
        for (int j = 0; j < ref->sgd->params[i]->size; ++j) {
            const double bump = 0.02 * (rand() / (RAND_MAX + 1.0) - 0.5) / std::sqrt((double)ref->sgd->params[i]->size);
            ref->sgd->params[i]->value[j] += bump;
            mine->sgd->params[i]->value[j] = ref->sgd->params[i]->value[j];
        }
    double worst_p = 0;
    for (size_t i = 0; i < mol.size(); ++i) {
        const double pr = ref->Predict(mol[i]), pm = mine->Predict(mol[i]);
        worst_p = std::max(worst_p, rel(pr, pm, std::fabs(pr)));
    }
    check("omega_Predict", worst_p, 1e-4);
    const double lr0 = ref->getLoss(4, &mol[0], targets), lm0 = mine->getLoss(4, &mol[0], targets);
    check("omega_getLoss", rel(lr0, lm0, lr0), 1e-4);
    double worst_l = 0;
    for (int e = 0; e < 3; ++e) {
        std::pair<double, double> a = ref->BatchLearn(4, &mol[0], targets, 0.001), b = mine->BatchLearn(4, &mol[0], targets, 0.001);
        worst_l = std::max(worst_l, std::max(rel(a.first, b.first, a.first), rel(a.second, b.second, a.second)));
    }
    check("omega_BatchLearn_losses", worst_l, 2e-4);
    double dq = 0, sq = 0;
    for (size_t i = 0; i < ref->sgd->params.size(); ++i)
        for (int j = 0; j < ref->sgd->params[i]->size; ++j) {
            dq = std::max(dq, std::fabs(ref->sgd->params[i]->value[j] - mine->sgd->params[i]->value[j]));
            sq = std::max(sq, std::fabs(ref->sgd->params[i]->value[j]));
        }
    check("omega_BatchLearn_parameters", dq / sq, 1e-3);
    mine->release();
}

// ::SMP_omega_physics (BASELINE config 3's model: halving channel widths, fields limited to max_field, every level feeds a
// hidden-layer read-out) vs ccn_b200::SMP_omega_physics: same seed, generic parameters, Predict / Feature / getLoss / three
// epochs of BatchLearn.
static void parity_physics(int L, int C, int max_field) {
    const int maxV = 12, F = 4, seed = 999;
    srand(seed);
    SMP_omega_physics *ref = new SMP_omega_physics(maxV, max_field, L, C, F);
    srand(seed);
    ccn_b200::SMP_omega_physics *mine = new ccn_b200::SMP_omega_physics(maxV, max_field, L, C, F);
    double dp = 0;
    if (ref->sgd->params.size() != mine->sgd->params.size()) dp = 1;
    for (size_t i = 0; i < ref->sgd->params.size() && dp == 0; ++i) {
        if (ref->sgd->params[i]->size != mine->sgd->params[i]->size) dp = 1;
        for (int j = 0; j < ref->sgd->params[i]->size && dp == 0; ++j)
            dp = std::max(dp, std::fabs(ref->sgd->params[i]->value[j] - mine->sgd->params[i]->value[j]));
    }
    check("physics_same_seed_parameters", dp, 0.0);
    std::vector<DenseGraph *> mol;
    srand(8);
    for (int i = 0; i < 4; ++i) mol.push_back(random_molecular_graph(8 + i));
    double targets[4] = {0.4, -0.7, 1.2, 0.1};
    for (size_t i = 0; i < ref->sgd->params.size(); ++i)
        for (int j = 0; j < ref->sgd->params[i]->size; ++j) {
            const double bump = 0.05 * (rand() / (RAND_MAX + 1.0) - 0.5) / std::sqrt((double)ref->sgd->params[i]->size);
            ref->sgd->params[i]->value[j] += bump;
            mine->sgd->params[i]->value[j] = ref->sgd->params[i]->value[j];
        }
    double worst_p = 0;
    for (size_t i = 0; i < mol.size(); ++i) {
        const double pr = ref->Predict(mol[i]), pm = mine->Predict(mol[i]);
        worst_p = std::max(worst_p, rel(pr, pm, std::fabs(pr)));
    }
    check("physics_Predict", worst_p, 1e-4);
    const double lr0 = ref->getLoss(4, &mol[0], targets), lm0 = mine->getLoss(4, &mol[0], targets);
    check("physics_getLoss", rel(lr0, lm0, lr0), 1e-4);
    double worst_l = 0;
    for (int e = 0; e < 3; ++e) {
        std::pair<double, double> a = ref->BatchLearn(4, &mol[0], targets, 0.001), b = mine->BatchLearn(4, &mol[0], targets, 0.001);
        worst_l = std::max(worst_l, std::max(rel(a.first, b.first, a.first), rel(a.second, b.second, a.second)));
    }
    check("physics_BatchLearn_losses", worst_l, 2e-4);
    double dq = 0, sq = 0;
    for (size_t i = 0; i < ref->sgd->params.size(); ++i)
        for (int j = 0; j < ref->sgd->params[i]->size; ++j) {
            dq = std::max(dq, std::fabs(ref->sgd->params[i]->value[j] - mine->sgd->params[i]->value[j]));
            sq = std::max(sq, std::fabs(ref->sgd->params[i]->value[j]));
        }
    check("physics_BatchLearn_parameters", dq / sq, 1e-3);
    mine->release();
}

// ::SMP_omega_pairgraphs vs ccn_b200::SMP_omega_pairgraphs: every example is a pair of graphs (the reference takes the second
// graph from the caller; two molecular graphs of different sizes here), same seed, generic parameters.
static void parity_pairgraphs(int L, int C, int max_field) {
    const int maxV = 12, F1 = 4, F2 = 4, seed = 31337;
    srand(seed);
    SMP_omega_pairgraphs *ref = new SMP_omega_pairgraphs(maxV, maxV, max_field, L, C, F1, F2);
    srand(seed);
    ccn_b200::SMP_omega_pairgraphs *mine = new ccn_b200::SMP_omega_pairgraphs(maxV, maxV, max_field, L, C, F1, F2);
    double dp = ref->sgd->params.size() == mine->sgd->params.size() ? 0 : 1;
    for (size_t i = 0; i < ref->sgd->params.size() && dp == 0; ++i) {
        if (ref->sgd->params[i]->size != mine->sgd->params[i]->size) dp = 1;
        for (int j = 0; j < ref->sgd->params[i]->size && dp == 0; ++j)
            dp = std::max(dp, std::fabs(ref->sgd->params[i]->value[j] - mine->sgd->params[i]->value[j]));
    }
    check("pairgraphs_same_seed_parameters", dp, 0.0);
    std::vector<DenseGraph *> m1, m2;
    srand(9);
    for (int i = 0; i < 3; ++i) {
        m1.push_back(random_molecular_graph(8 + i));
        m2.push_back(random_molecular_graph(10 - i));
    }
    double targets[3] = {0.4, -0.7, 1.2};
    for (size_t i = 0; i < ref->sgd->params.size(); ++i)
        for (int j = 0; j < ref->sgd->params[i]->size; ++j) {
            const double bump = 0.05 * (rand() / (RAND_MAX + 1.0) - 0.5) / std::sqrt((double)ref->sgd->params[i]->size);
            ref->sgd->params[i]->value[j] += bump;
            mine->sgd->params[i]->value[j] = ref->sgd->params[i]->value[j];
        }
    double worst_p = 0;
    for (size_t i = 0; i < m1.size(); ++i) {
        const double pr = ref->Predict(m1[i], m2[i]), pm = mine->Predict(m1[i], m2[i]);
        worst_p = std::max(worst_p, rel(pr, pm, std::fabs(pr)));
    }
    check("pairgraphs_Predict", worst_p, 1e-4);
    const double lr0 = ref->getLoss(3, &m1[0], &m2[0], targets), lm0 = mine->getLoss(3, &m1[0], &m2[0], targets);
    check("pairgraphs_getLoss", rel(lr0, lm0, lr0), 1e-4);
    double worst_l = 0;
    for (int e = 0; e < 3; ++e) {
        std::pair<double, double> a = ref->BatchLearn(3, &m1[0], &m2[0], targets, 0.001), b = mine->BatchLearn(3, &m1[0], &m2[0], targets, 0.001);
        worst_l = std::max(worst_l, std::max(rel(a.first, b.first, a.first), rel(a.second, b.second, a.second)));
    }
    check("pairgraphs_BatchLearn_losses", worst_l, 2e-4);
    double dq = 0, sq = 0;
    for (size_t i = 0; i < ref->sgd->params.size(); ++i)
        for (int j = 0; j < ref->sgd->params[i]->size; ++j) {
            dq = std::max(dq, std::fabs(ref->sgd->params[i]->value[j] - mine->sgd->params[i]->value[j]));
            sq = std::max(sq, std::fabs(ref->sgd->params[i]->value[j]));
        }
    check("pairgraphs_BatchLearn_parameters", dq / sq, 1e-3);
    mine->release();
}

// ::SMP_2D_ver8 (K_l stored [C, 18 C], Momentum) vs ccn_b200::SMP_2D_ver8: same seed, generic parameters, Predict / getLoss /
// gradients / three epochs of BatchLearn with the reference's Momentum object on both sides.
static void parity_ver8(int L, int C, int D) {
    const int maxV = 12, F = 4, seed = 4242;
    srand(seed);
    SMP_2D_ver8 *ref = new SMP_2D_ver8(maxV, L, C, F, D, 0.9);
    srand(seed);
    ccn_b200::SMP_2D_ver8 *mine = new ccn_b200::SMP_2D_ver8(maxV, L, C, F, D, 0.9);
    double dp = 0;
    for (size_t i = 0; i < ref->sgd->params.size(); ++i)
        for (int j = 0; j < ref->sgd->params[i]->size; ++j)
            dp = std::max(dp, std::fabs(ref->sgd->params[i]->value[j] - mine->sgd->params[i]->value[j]));
    check("ver8_same_seed_parameters", dp, 0.0);
    std::vector<DenseGraph *> mol;
    srand(6);
    for (int i = 0; i < 4; ++i) mol.push_back(random_molecular_graph(7 + i));
    double targets[4] = {0.4, -0.7, 1.2, 0.1};
    for (size_t i = 0; i < ref->sgd->params.size(); ++i)
        for (int j = 0; j < ref->sgd->params[i]->size; ++j) {
            const double bump = 0.02 * (rand() / (RAND_MAX + 1.0) - 0.5) / std::sqrt((double)ref->sgd->params[i]->size);
            ref->sgd->params[i]->value[j] += bump;
            mine->sgd->params[i]->value[j] = ref->sgd->params[i]->value[j];
        }
    double worst_p = 0;
    for (size_t i = 0; i < mol.size(); ++i) {
        const double pr = ref->Predict(mol[i]), pm = mine->Predict(mol[i]);
        worst_p = std::max(worst_p, rel(pr, pm, std::fabs(pr)));
    }
    check("ver8_Predict", worst_p, 1e-4);
    double worst_l = 0;
    for (int e = 0; e < 3; ++e) {
        std::pair<double, double> a = ref->BatchLearn(4, &mol[0], targets, 0.001), b = mine->BatchLearn(4, &mol[0], targets, 0.001);
        worst_l = std::max(worst_l, std::max(rel(a.first, b.first, a.first), rel(a.second, b.second, a.second)));
    }
    check("ver8_BatchLearn_losses", worst_l, 2e-4);
    double dq = 0, sq = 0;
    for (size_t i = 0; i < ref->sgd->params.size(); ++i)
        for (int j = 0; j < ref->sgd->params[i]->size; ++j) {
            dq = std::max(dq, std::fabs(ref->sgd->params[i]->value[j] - mine->sgd->params[i]->value[j]));
            sq = std::max(sq, std::fabs(ref->sgd->params[i]->value[j]));
        }
    check("ver8_BatchLearn_parameters", dq / sq, 1e-3);
    mine->release();
}

static DenseGraph *random_molecular_graph(int V) {
    DenseGraph *g = new DenseGraph(V, 4);
    std::vector<int> deg(V, 0);
    for (int v = 1; v < V; ++v) {  // random spanning tree, degree <= 4
        int u = rand() % v, tries = 0;
        while (deg[u] >= 4 && tries++ < 64) u = rand() % v;
        g->adj[u][v] = g->adj[v][u] = 1;
        ++deg[u];
        ++deg[v];
    }
    for (int k = 0; k < V / 8; ++k) {
        const int u = rand() % V, v = rand() % V;
        if (u != v && !g->adj[u][v] && deg[u] < 4 && deg[v] < 4) {
            g->adj[u][v] = g->adj[v][u] = 1;
            ++deg[u];
            ++deg[v];
        }
    }
    for (int v = 0; v < V; ++v) g->feature[v][rand() % 4] = 1.0;
    return g;
}

static void bench(int L, int C, int V, int B) {
    srand(99);
    ccn_b200::SMP_beta *net = new ccn_b200::SMP_beta(V, L, C, 4, 2);
    std::vector<DenseGraph *> graphs(B);
    std::vector<double> targets(B);
    long long contractions = 0;
    for (int i = 0; i < B; ++i) {
        graphs[i] = random_molecular_graph(V);
        targets[i] = (rand() % 100) / 50.0 - 1.0;
        contractions += (long long)V * L;
    }
    net->BatchLearn(B, &graphs[0], &targets[0], 1e-4);  // warm-up: graph tables, device buffers
    double best = 1e30;
    for (int r = 0; r < 3; ++r) {
        const auto t0 = std::chrono::steady_clock::now();
        net->BatchLearn(B, &graphs[0], &targets[0], 1e-4);
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (s < best) best = s;
    }
    double best_g = 1e30, best_l = 1e30;  // the two device passes of a BatchLearn on their own
    for (int r = 0; r < 3; ++r) {
        auto t0 = std::chrono::steady_clock::now();
        net->Gradients(B, &graphs[0], &targets[0]);
        best_g = std::min(best_g, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        t0 = std::chrono::steady_clock::now();
        net->getLoss(B, &graphs[0], &targets[0]);
        best_l = std::min(best_l, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    }
    std::fprintf(stderr, "model bench: Gradients (forward + backward) %.3f ms, getLoss (forward) %.3f ms\n", best_g * 1e3, best_l * 1e3);
    // BatchLearn = forward+backward (with gradients) + optimizer + one more forward (the "loss after"): count its contractions once
    std::printf("{\"what\": \"ccn_b200::SMP_beta::BatchLearn (C++ model facade, host DenseGraph inputs): graph tables cached, level 0, "
                "%d fused-promotion levels, read-out, loss, gradients, host Adam, second forward for the loss after\", "
                "\"graphs\": %d, \"vertices\": %d, \"levels\": %d, \"channels\": %d, \"s_per_BatchLearn\": %.6f, "
                "\"graphs_per_s\": %.1f, \"value\": %.1f, \"unit\": \"contractions/s (fwd+bwd; real ragged receptive fields)\"}\n",
                L, B, V, L, C, best, B / best, contractions / best);
    net->release();
}

int main(int argc, char **argv) {
    const std::string what = argc > 1 ? argv[1] : "parity";
    const int a1 = argc > 2 ? std::atoi(argv[2]) : 0, a2 = argc > 3 ? std::atoi(argv[3]) : 0, a3 = argc > 4 ? std::atoi(argv[4]) : 0,
              a4 = argc > 5 ? std::atoi(argv[5]) : 0;
    if (what == "bench") {
        bench(a1 ? a1 : 3, a2 ? a2 : 32, a3 ? a3 : 24, a4 ? a4 : 128);
        return 0;
    }
    if (a1) {
        parity(a1, a2, a3);
    } else {
        parity(1, 10, 5);  // the reference test's own configuration (generic kernels, SIMT mix)
        parity(2, 8, 2);   // fused kernels (C = 8), two levels
        parity(2, 32, 2);  // fused kernels + tensor-core mix
        parity_omega(2, 8, 2, 5);   // SMP_omega: fields of the 8..11-vertex graphs cut to 5 members
        parity_omega(3, 16, 1, 6);
        parity_physics(2, 8, 5);    // SMP_omega_physics: widths 8 -> 4 -> 2, multi-level hidden-layer read-out
        parity_physics(3, 64, 6);   //   64 -> 32 -> 16 -> 8: fused kernels and tensor-core mix on the wide levels
        parity_pairgraphs(2, 8, 5); // SMP_omega_pairgraphs: two trunks, level-wise concatenation, two hidden layers
        parity_pairgraphs(2, 32, 6);
        parity_ver8(2, 8, 2);       // SMP_2D_ver8: K stored transposed, Momentum
        parity_ver8(2, 32, 1);
    }
    std::printf("model failures=%d\n", failures);
    return failures == 0 ? 0 : 1;
}
