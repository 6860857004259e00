// compile-only check of the drop-in spelling: a translation unit written against GraphFlow_gpu's class names
#include "Matrix.h"
#include "Tensor4D.h"
#define CCN_B200_DROP_IN
#include "graphflow_b200/ccn_ops_b200.h"
int main() {
    Tensor4D *T = new Tensor4D(4, 4, 4, 2);
    Matrix *adj = new Matrix(4, 4);
    RisiContraction_18_gpu *op = new RisiContraction_18_gpu(T, adj);   // the reference's class name
    Matrix *A = new Matrix(3, 5), *B = new Matrix(5, 2);
    MatMul_gpu *mm = new MatMul_gpu(A, B);
    (void)op; (void)mm;
    return 0;
}
