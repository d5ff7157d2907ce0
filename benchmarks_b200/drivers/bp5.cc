// bp5.cc -- the reference's bp5_kokkos driver protocol (bp5_kokkos/benchmark.cc) on the B200 operator.
//
//   bp5 <degree> [s] [compact]
//
// Same mesh family (create_triangulation.h:17-29: subdivisions (1|2, 1|2, 4 + s%4), box [0, subdivisions],
// refine_global(s / 12)), same operator -- Helmholtz (grad v, grad u) + (v, u) with QGauss(p+1)
// (benchmark.cc:62-137, 180-191) -- Jacobi-preconditioned CG capped at 100 iterations with
// ReductionControl(100, 1e-15, 1e-8), NoConvergence swallowed (:355-378), rhs[i] = i % 8 on the unconstrained
// DoFs (:341-347), best of 4 solves and of 4 batches of 50 operator applications (:386-395), and the
// reference's one-line table row (:400-409).  Single process / single GPU.
#include <b200fe/operator.hpp>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>

using namespace b200fe;
using clk = std::chrono::steady_clock;
static double since(clk::time_point t0) { return std::chrono::duration<double>(clk::now() - t0).count(); }

template <int fe_degree>
void test(const unsigned int s, const bool short_output)
{
    constexpr int n_q_points = fe_degree + 1;
    auto t0 = clk::now();
    const unsigned int n_refine = s / 12, remainder = s % 12;
    int sub[3] = {1, 1, (int)(4 + remainder % 4)};
    if (remainder > 3) sub[1] = 2;
    if (remainder > 7) sub[0] = 2;
    const double p1[3] = {0, 0, 0}, p2[3] = {(double)sub[0], (double)sub[1], (double)sub[2]};
    BoxMesh mesh(sub, (int)n_refine, fe_degree, p1, p2, 1, 0, B200FE_PARTITION_BLOCKS);
    // B200FE_GEOMETRY=onthefly: the mesh is a box of unit cubes -> separable Helmholtz kernel, no G / JxW in memory
    const char *geo = std::getenv("B200FE_GEOMETRY");
    const Geometry geometry = geo && std::string(geo) == "onthefly" ? Geometry::OnTheFly : Geometry::Stored;
    LaplaceOperator<3, fe_degree, n_q_points, double> helmholtz(mesh, Quadrature::Gauss, B200FE_OP_HELMHOLTZ, 1, {}, geometry);
    Vector solution, rhs;
    helmholtz.initialize_dof_vector(solution);
    rhs.reinit(solution);
    helmholtz.compute_diagonal();
    const double *preconditioner = helmholtz.get_matrix_diagonal_inverse();
    {
        std::vector<double> h(mesh.info.n_owned);
        std::vector<uint32_t> con(mesh.info.n_constrained);
        check(b200fe_boxmesh_fill(mesh.handle(), nullptr, con.data(), nullptr, nullptr, nullptr, nullptr));
        for (size_t i = 0; i < h.size(); ++i) h[i] = (double)(i % 8);
        for (uint32_t c : con) h[c] = 0.0;
        rhs.from_host(h);
    }
    cudaDeviceSynchronize();
    if (!short_output) std::printf("Setup time:         %g s\n", since(t0));

    ReductionControl solver_control(100, 1e-15, 1e-8);
    SolverCG solver(solver_control, /*check_every=*/4);
    double solver_time = 1e10;
    for (unsigned int t = 0; t < 4; ++t) {
        solution = 0;
        auto tt = clk::now();
        try {
            solver.solve(helmholtz, solution, rhs, preconditioner);
        } catch (NoConvergence &) {
            // prevent the solver to throw an exception in case we should need more than 100 iterations
        }
        cudaDeviceSynchronize();
        solver_time = std::min(since(tt), solver_time);
    }
    double matvec_time = 1e10;
    for (unsigned int t = 0; t < 4; ++t) {
        auto tt = clk::now();
        for (unsigned int i = 0; i < 50; ++i) helmholtz.vmult(rhs, solution);
        cudaDeviceSynchronize();
        matvec_time = std::min(since(tt) / 50, matvec_time);
    }
    std::printf("%2d | %2d |%10llu |%11llu | %11.4e | %11.4e | %4u | %11.4e\n", fe_degree, n_q_points, mesh.n_global_active_cells(),
                mesh.n_dofs(), solver_time / solver_control.last_step(), mesh.n_dofs() / solver_time * solver_control.last_step(),
                solver_control.last_step(), matvec_time);
}

template <int fe_degree>
void do_test(const int s_in, const bool compact_output)
{
    if (s_in < 1) {
        unsigned int s = std::max(3U, static_cast<unsigned int>(std::log2(1024 / fe_degree / fe_degree / fe_degree)));
        std::cout << " p |  q | n_element |     n_dofs |     time/it |   dofs/s/it | itCG | time/matvec" << std::endl;
        unsigned long long p3 = 1;
        for (int d = 0; d < 3; ++d) p3 *= (fe_degree + 1);
        while ((2 + p3) * (1ULL << (s / 4)) < 12000000ULL) {
            test<fe_degree>(s, compact_output);
            ++s;
        }
        std::cout << std::endl << std::endl;
    } else
        test<fe_degree>(s_in, compact_output);
}

int main(int argc, char **argv)
{
    try {
        unsigned int degree = 1;
        int s = -1;
        bool compact_output = true;
        if (argc > 1) degree = std::atoi(argv[1]);
        if (argc > 2) s = std::atoi(argv[2]);
        if (argc > 3) compact_output = std::atoi(argv[3]);
        switch (degree) {
            case 1: do_test<1>(s, compact_output); break;
            case 2: do_test<2>(s, compact_output); break;
            case 3: do_test<3>(s, compact_output); break;
            case 4: do_test<4>(s, compact_output); break;
            case 5: do_test<5>(s, compact_output); break;
            case 6: do_test<6>(s, compact_output); break;
            case 7: do_test<7>(s, compact_output); break;
            case 8: do_test<8>(s, compact_output); break;
            default: std::cout << "Degree " << degree << " not implemented" << std::endl;
        }
    } catch (std::exception &e) {
        std::cerr << e.what() << std::endl;
        return 1;
    }
    return 0;
}
