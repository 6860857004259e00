// Compile-only check (tests/cpp/Makefile, -fsyntax-only, 32-bit reference tree): with CCN_B200_DROP_IN the model header gives
// the reference's spelling `SMP_beta`, and the body of the reference's tests/test_SMP_beta.cpp (BatchLearn with float targets,
// Predict, save_model, load_model) type-checks against it.
#include <string>

#include "graphflow_b200/SMP_beta_b200.h"

int dropin_model_check() {
    SMP_beta train_network(10, 1, 10, 4, 5), test_network(10, 1, 10, 4, 5);
    DenseGraph *graphs[1] = {new DenseGraph(3, 4)};
    float targets[1] = {1.0f};
    train_network.BatchLearn(1, graphs, targets, 0.001f);
    float predict = train_network.Predict(graphs[0]);
    train_network.save_model(std::string("m.dat"));
    test_network.load_model(std::string("m.dat"));
    // the other facades under their reference names (SMP_omega.h:32, SMP_omega_physics.h:31, SMP_2D_ver8.h:32)
    SMP_omega omega(10, 5, 2, 8, 4, 2);
    SMP_omega_physics physics(10, 5, 2, 8, 4);
    SMP_2D_ver8 ver8(10, 2, 8, 4, 2, 0.9);
    double t[1] = {1.0};
    physics.BatchLearn(1, graphs, t, 0.001);
    SMP_omega_pairgraphs pairs(10, 10, 3, 2, 8, 4, 4);  // SMP_omega_pairgraphs.h:53
    pairs.BatchLearn(1, graphs, graphs, t, 0.001);
    predict += (float)pairs.Predict(graphs[0], graphs[0]);
    return predict > 0 && omega.Predict(graphs[0]) + physics.Predict(graphs[0]) + ver8.Predict(graphs[0]) > 0;
}
