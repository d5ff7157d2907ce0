// bp6.cc -- CEED BP6 (vector Laplacian, 3 components, collocated Gauss-Lobatto quadrature nq = p + 1) on a smoothly
// deformed box mesh whose first octant is refined once more (hanging nodes): BASELINE config C5.
//
//   bp6 <degree> [minsize] [maxsize] [hanging=1] [deform_amplitude=0.05]
//
// The reference has no BP6 / hanging-node / deformed-mesh program; this driver follows the
// protocol of its BP3 driver (CEED_bp/src/bp3.cc): same box sweep (:433-488) -- here with the cells of the lower
// octant refined once --, right-hand side int phi_i in every component (:184-239), CG with
// ReductionControl(1e9, 1e-16, 1e-9) from x = 0, best of 3 solves (:266-288), best of 5 batches of operator
// applications (:290-314).  Sizes count the DoFs of all three components.  One line per mesh:
//   cells | dofs(3 comp.) | hanging rows | matvec s | CG s | CG s/it | its | GDoF/s apply | GDoF/s CG
// Single process / single GPU (multi-GPU runs go through the Python launcher).
#include <b200fe/operator.hpp>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <memory>

using namespace b200fe;
using clk = std::chrono::steady_clock;
static double since(clk::time_point t0) { return std::chrono::duration<double>(clk::now() - t0).count(); }

template <int fe_degree>
void run(std::size_t min_size, std::size_t max_size, bool hanging, double amplitude)
{
    constexpr int nq = fe_degree + 1, n_comp = 3;
    std::printf("Testing FE_Q<3>(%d)^3, n_q_points_1d = %d (QGaussLobatto), %s, deformation amplitude %g\nNo. of GPUs: 1\n", fe_degree, nq,
                hanging ? "lower octant refined once (hanging nodes)" : "uniform mesh", amplitude);
    std::printf("  cells |      dofs | hanging |   matvec |  CG_time | CG_time/it |  its | apply GDoF/s | CG GDoF/s\n");
    for (unsigned cycle = 3; cycle < 38; ++cycle) {  // from 2 x 2 x 2 cells on: the octant is at least one cell
        const unsigned n_refine = cycle / 3, rem = cycle % 3;
        int sub[3], lo[3] = {0, 0, 0}, hi[3];
        double p1[3] = {-1, -1, -1}, p2[3];
        for (int d = 0; d < 3; ++d) {
            sub[d] = d < (int)rem ? 2 : 1;
            p2[d] = d < (int)rem ? 2.8 : 0.9;
            hi[d] = (sub[d] << n_refine) / 2;
        }
        // projected size: uniform part + the extra DoFs of the refined octant, three components
        std::size_t projected = n_comp * BoxMesh::bp3_projected_size(cycle, fe_degree);
        if (hanging) projected += n_comp * (std::size_t)(7 * BoxMesh::bp3_projected_size(cycle, fe_degree) / 8);
        if (projected < min_size) continue;
        if (projected > max_size) { std::printf("Projected size %zu higher than max size, terminating.\n", projected); break; }
        const Deformation deform{amplitude, 2.0};
        unsigned long long cells = 0, dofs = 0, n_hang = 0;
        std::unique_ptr<LaplaceOperator<3, fe_degree, nq, double>> A;
        if (hanging) {
            HangingBoxMesh mesh(sub, (int)n_refine, fe_degree, lo, hi, p1, p2);
            cells = mesh.n_global_active_cells(); dofs = mesh.n_dofs(); n_hang = mesh.info.n_hanging_rows;
            A = std::make_unique<LaplaceOperator<3, fe_degree, nq, double>>(mesh, Quadrature::GaussLobatto, B200FE_OP_LAPLACE, 2, deform);
        } else {
            BoxMesh mesh(sub, (int)n_refine, fe_degree, p1, p2);
            cells = mesh.n_global_active_cells(); dofs = mesh.n_dofs();
            A = std::make_unique<LaplaceOperator<3, fe_degree, nq, double>>(mesh, Quadrature::GaussLobatto, B200FE_OP_LAPLACE, 2, deform);
        }
        Vector solution, rhs;
        A->initialize_dof_vector(solution, n_comp);
        rhs.reinit(solution);
        A->compute_rhs(rhs);
        cudaStream_t solve_stream;
        check_cuda(cudaStreamCreate(&solve_stream), "cudaStreamCreate");
        double time_cg = 1e10;
        unsigned its = 0;
        for (unsigned i = 0; i < 3; ++i) {
            ReductionControl solver_control(1000000000, 1e-16, 1e-9);
            SolverCG cg(solver_control);
            cg.set_stream(solve_stream);
            solution = 0;
            cudaDeviceSynchronize();
            auto t = clk::now();
            cg.solve(*A, solution, rhs, PreconditionIdentity());
            cudaDeviceSynchronize();
            time_cg = std::min(time_cg, since(t));
            its = solver_control.last_step();
        }
        cudaStreamDestroy(solve_stream);
        A->distribute(solution);
        const unsigned n_mv = n_comp * dofs < 10000000 ? 200 : 50;
        double mv = 1e10;
        for (unsigned i = 0; i < 5; ++i) {
            cudaDeviceSynchronize();
            auto t = clk::now();
            for (unsigned k = 0; k < n_mv; ++k) A->vmult(solution, rhs);
            cudaDeviceSynchronize();
            mv = std::min(mv, since(t) / n_mv);
        }
        const double nd = (double)n_comp * dofs;
        std::printf("%7llu | %9llu | %7llu | %.2e | %.2e |   %.2e | %4u | %12.3f | %9.3f\n", cells, (unsigned long long)(n_comp * dofs), n_hang, mv,
                    time_cg, time_cg / its, its, 1e-9 * nd / mv, 1e-9 * nd * its / time_cg);
    }
}

int main(int argc, char **argv)
{
    try {
        if (argc == 1) { std::cout << "Expected at least one argument.\nUsage:\n./bp6 degree [minsize] [maxsize] [hanging=1] [deform_amplitude=0.05]\n"; return 1; }
        const unsigned degree = std::atoi(argv[1]);
        const std::size_t minsize = argc > 2 ? std::atoll(argv[2]) : 1, maxsize = argc > 3 ? std::atoll(argv[3]) : static_cast<std::size_t>(-1);
        const bool hanging = argc > 4 ? std::atoi(argv[4]) != 0 : true;
        const double amplitude = argc > 5 ? std::atof(argv[5]) : 0.05;
        switch (degree) {
            case 1: run<1>(minsize, maxsize, hanging, amplitude); break;
            case 2: run<2>(minsize, maxsize, hanging, amplitude); break;
            case 3: run<3>(minsize, maxsize, hanging, amplitude); break;
            case 4: run<4>(minsize, maxsize, hanging, amplitude); break;
            case 5: run<5>(minsize, maxsize, hanging, amplitude); break;
            case 6: run<6>(minsize, maxsize, hanging, amplitude); break;
            case 7: run<7>(minsize, maxsize, hanging, amplitude); break;
            case 8: run<8>(minsize, maxsize, hanging, amplitude); break;
            default: std::cerr << "degree outside 1..8: no work" << std::endl;
        }
    } catch (std::exception &exc) {
        std::cerr << "\n\n----------------------------------------------------\nException on processing: \n"
                  << exc.what() << "\nAborting!\n----------------------------------------------------" << std::endl;
        return 1;
    }
    return 0;
}
