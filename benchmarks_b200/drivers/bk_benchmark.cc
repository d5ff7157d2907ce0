// bk_benchmark.cc -- the reference's standalone kernel drivers (CEED_BK/src/BK{1,3,5}/templated_cuda_benchmark.cc)
// on the B200 kernels, FP64.
//
//   bk_benchmark <bk1|bk3|bk5> [p=2] [nelmt=2<<18] [ntests=10]
//
// Same seedless synthetic inputs (in = 3, JxW = 1, G = 2, basis[q*nm+i] = cos(q*nm+i), dbasis[i*nq+n] =
// cos(i*nq+n), :45-66), same timing (host chrono around launch + cudaDeviceSynchronize, min over ntests,
// :92-101), same metric formulas (GDOF/s = nelmt nm^3 / t, bw = (2 nDOF + 6 nQuad) sizeof(T) / t, :111-114)
// and the same table columns as CEED_BK/include/benchmark_printer.hpp.
#include <b200fe/operator.hpp>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <string>

using namespace b200fe;

int main(int argc, char **argv)
{
    try {
        const std::string kind = argc > 1 ? argv[1] : "bk3";
        const unsigned p = argc > 2 ? std::atoi(argv[2]) : 2u;
        const unsigned nelmt = argc > 3 ? std::atoi(argv[3]) : 2u << 18;
        const unsigned ntests = argc > 4 ? std::atoi(argv[4]) : 10u;
        const int k = kind == "bk1" ? 1 : kind == "bk3" ? 3 : kind == "bk5" ? 5 : 0;
        if (!k) { std::fprintf(stderr, "usage: bk_benchmark <bk1|bk3|bk5> [p] [nelmt] [ntests]\n"); return 1; }
        const unsigned nm = p + 1, nq = k == 5 ? p + 1 : p + 2;
        const size_t nDOF = (size_t)nm * nm * nm * nelmt, nQuad = (size_t)nq * nq * nq * nelmt;
        std::vector<double> basis(nq * nm), dbasis(nq * nq);
        for (unsigned i = 0; i < nq * nm; ++i) basis[i] = std::cos((double)i);
        for (unsigned i = 0; i < nq * nq; ++i) dbasis[i] = std::cos((double)i);
        DeviceArray<double> d_in, d_out(nDOF), d_geo, d_sum(1);
        { std::vector<double> h(nDOF, 3.0); d_in.upload(h.data(), h.size()); }
        { std::vector<double> h(k == 1 ? nQuad : 6 * nQuad, k == 1 ? 1.0 : 2.0); d_geo.upload(h.data(), h.size()); }
        auto launch = [&]() {
            if (k == 1) check(b200fe_bk1_apply(p, nq, nelmt, basis.data(), d_geo.data(), d_in.data(), d_out.data(), nullptr));
            else if (k == 3) check(b200fe_bk3_apply(p, nq, nelmt, basis.data(), dbasis.data(), d_geo.data(), d_in.data(), d_out.data(), nullptr));
            else check(b200fe_bk5_apply(p, nelmt, dbasis.data(), d_geo.data(), d_in.data(), d_out.data(), nullptr));
        };
        double time = std::numeric_limits<double>::max();
        for (unsigned t = 0; t < ntests; ++t) {
            auto t0 = std::chrono::steady_clock::now();
            launch();
            check_cuda(cudaDeviceSynchronize(), "sync");
            time = std::min(time, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        }
        check(b200fe_sum_squares(nDOF, d_out.data(), d_sum.data(), nullptr));
        double sum = 0;
        d_sum.download(&sum);
        int epb = 0, blocks = 0, threads = 0, smem = 0;
        check(b200fe_bk_launch_info(k, p, nq, nelmt, &epb, &blocks, &threads, &smem));
        const double gdofs = 1e-9 * nDOF / time;
        const double bw = 1e-9 * (2.0 * nDOF + (k == 1 ? 1.0 : 6.0) * nQuad) * sizeof(double) / time;
        std::printf("%8s %3s %10s %14s %10s %16s %12s %12s %10s %10s %14s\n", "Kernel", "p", "nelmt", "nelmtPerBatch", "numBlocks",
                    "threadsPerBlock", "DOF", "time", "GDOF/s", "bw(GB/s)", "check");
        std::printf("%8s %3u %10u %14d %10d %16d %12zu %12.4e %10.4f %10.2f %14.8g\n", kind == "bk1" ? "BK1" : kind == "bk3" ? "BK3" : "BK5", p,
                    nelmt, epb, blocks, threads, nDOF, time, gdofs, bw, std::sqrt(sum));
    } catch (std::exception &e) {
        std::fprintf(stderr, "%s\n", e.what());
        return EXIT_FAILURE;
    }
    return 0;
}
