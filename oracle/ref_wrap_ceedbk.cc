// oracle/ref_wrap_ceedbk.cc -- TEST INFRASTRUCTURE ONLY.
// extern "C" shim around the reference's OWN serial kernels, compiled in place from
// /root/reference/CEED_BK/include (never copied into this repo).  Output: oracle/_ref/libref_ceedbk.so.
//   BK1::Serial::SumFactorization  CEED_BK/include/kernels/BK1/serial_kernels.hpp:9-135
//   BK3::Serial::SumFactorization  CEED_BK/include/kernels/BK3/serial_kernels.hpp:10-197
//   BK5::Serial::SumFactorization  CEED_BK/include/kernels/BK5/serial_kernels.hpp:10-85  (reads G comp. 0 only)
#include <algorithm>
#include <cmath>
#include <iostream>
#include <kernels/BK1/serial_kernels.hpp>
#include <kernels/BK3/serial_kernels.hpp>
#include <kernels/BK5/serial_kernels.hpp>

extern "C" {
double ref_ceedbk_bk1(unsigned nq, unsigned nelmt, const double *basis, const double *JxW, double *in, double *out)
{
    return BK1::Serial::SumFactorization<double>(nq, nq, nq, nelmt, basis, basis, basis, JxW, in, out);
}
double ref_ceedbk_bk3(unsigned nq, unsigned nelmt, const double *basis, const double *dbasis, const double *G,
                      double *in, double *out)
{
    return BK3::Serial::SumFactorization<double>(nq, nq, nq, nelmt, basis, basis, basis, dbasis, dbasis, dbasis, G, in, out);
}
double ref_ceedbk_bk5(unsigned nq, unsigned nelmt, const double *dbasis, const double *G, const double *in, double *out)
{
    return BK5::Serial::SumFactorization<double>(nq, nq, nq, nelmt, dbasis, dbasis, dbasis, G, in, out);
}
}
