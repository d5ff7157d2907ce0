// oracle/ref_wrap_sumfact.cc -- TEST INFRASTRUCTURE ONLY.
// extern "C" shim around the reference's serial kernels of the sum_factorization project, compiled in
// place from /root/reference/sum_factorization/include.  Output: oracle/_ref/libref_sumfact.so.
//   BK1::Serial::DirectEvaluation  sum_factorization/include/kernels/BK1/serial_kernels.hpp:9-61
//   BK1::Serial::SumFactorization  sum_factorization/include/kernels/BK1/serial_kernels.hpp:66-
//   BK5::Serial::SumFactorization  sum_factorization/include/kernels/BK5/serial_kernels.hpp:10-85 (G [e][i][j][6][k])
#include <algorithm>
#include <cmath>
#include <iostream>
#include <kernels/BK1/serial_kernels.hpp>
#include <kernels/BK5/serial_kernels.hpp>

extern "C" {
double ref_sumfact_bk1_direct(unsigned nq, unsigned nelmt, const double *basis, const double *JxW, double *in, double *out)
{
    return BK1::Serial::DirectEvaluation<double>(nq, nq, nq, nelmt, basis, basis, basis, JxW, in, out);
}
double ref_sumfact_bk1(unsigned nq, unsigned nelmt, const double *basis, const double *JxW, double *in, double *out)
{
    return BK1::Serial::SumFactorization<double>(nq, nq, nq, nelmt, basis, basis, basis, JxW, in, out);
}
double ref_sumfact_bk5(unsigned nq, unsigned nelmt, const double *dbasis, const double *G, const double *in, double *out)
{
    return BK5::Serial::SumFactorization<double>(nq, nq, nq, nelmt, dbasis, dbasis, dbasis, G, in, out);
}
}
