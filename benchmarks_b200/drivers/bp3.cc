// bp3.cc -- the reference's BP3 driver protocol (CEED_bp/src/bp3.cc) on the B200 operator.
//
//   bp3 <degree> [minsize] [maxsize] [nq_offset=2] [quad=gauss|gll]
//
// Same mesh sweep (bp3.cc:433-488), same right-hand side (int phi, :184-239), same solver protocol:
// 10x CG with ReductionControl(1e9, 1e-16, 1e-9) from x = 0, best time (:266-288); 5 batches of
// 200 (< 1e7 DoFs) or 50 operator applications, best batch (:290-314); vmult_dummy timings with
// ghost exchange / computation switched (:331-408).  Output: the reference's two tables
// (cells, dofs, matvec, CG_tot_time, CG_time/iters, cg_its, cg_reduction) and
// (cells, dofs, mv_ghost_and_compute, mv_compute_only, mv_ghost_only), plus GDoF/s.
// One process per GPU: rank / size are taken from the launcher's environment (e.g. `torchrun --no-python --nproc-per-node N
// ./bp3 4 ...`; b200fe::Communicator), the mesh is partitioned along the p4est curve like the reference's
// parallel::distributed::Triangulation (bp3.cc:83), times are the maximum over ranks (:301-302).  The N > 1 path of this
// driver has not been run yet (round 1 measured multi-GPU through the Python launcher); N = 1 is the tested path.
#include <b200fe/operator.hpp>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <memory>

using namespace b200fe;
using clk = std::chrono::steady_clock;
static double since(clk::time_point t0) { return std::chrono::duration<double>(clk::now() - t0).count(); }

template <int fe_degree, int nq>
struct LaplaceProblem {
    struct Row { unsigned long long cells, dofs; double mv, cg_tot, cg_per_it; unsigned its; double red, both, comp, ghost; };
    std::vector<Row> table;

    void run(std::size_t min_size, std::size_t max_size, Quadrature quad)
    {
        const Communicator comm = Communicator::from_environment();
        check_cuda(cudaSetDevice(comm.local_rank), "cudaSetDevice");
        const bool root = comm.rank == 0;
        // the tables are printed by rank 0 only (pcout of bp3.cc:92); the other ranks keep a valid stdout on /dev/null
        if (!root && !std::freopen("/dev/null", "w", stdout)) throw Error(B200FE_ERR_INVALID_ARG, "cannot redirect stdout");
        std::printf("Testing FE_Q<3>(%d), n_q_points_1d = %d (%s)\nNo. of GPUs: %d\n", fe_degree, nq, quad == Quadrature::Gauss ? "QGauss" : "QGaussLobatto", comm.size);
        for (unsigned cycle = 0; cycle < 38; ++cycle) {
            const std::size_t projected = BoxMesh::bp3_projected_size(cycle, fe_degree);
            if (projected < min_size) continue;
            if (projected > max_size) { std::printf("Projected size %zu higher than max size, terminating.\n\n", projected); break; }
            std::printf("Cycle %u\n", cycle);
            auto t0 = clk::now();
            BoxMesh mesh = BoxMesh::bp3_cycle(cycle, fe_degree, comm.size, comm.rank);
            std::printf("  Number of cells: %llu |   Number of DoFs: %llu\n", mesh.n_global_active_cells(), mesh.n_dofs());
            // B200FE_GEOMETRY=onthefly: geometric factors rebuilt in the kernel instead of streamed; the
            // reference's meshes are axis-aligned cubes, so "bp3 <p> <min> <max> 1 gll" then runs the separable kernel
            const char *geo = std::getenv("B200FE_GEOMETRY");
            const Geometry geometry = geo && std::string(geo) == "onthefly" ? Geometry::OnTheFly : Geometry::Stored;
            LaplaceOperator<3, fe_degree, nq, double> system_matrix(mesh, quad, B200FE_OP_LAPLACE, 1, {}, geometry);
            if (geometry == Geometry::OnTheFly) {
                int cart = 0;
                check(b200fe_op_cartesian(system_matrix.handle(), &cart));
                std::printf("  Geometry on the fly (%s kernel)\n", cart ? "cartesian" : "affine");
            }
            std::unique_ptr<Halo> halo;
            if (comm.size > 1) {
                halo = std::make_unique<Halo>(mesh, comm);
                system_matrix.set_halo(*halo);
            }
            auto max_over_ranks = [&](double t) { return halo ? halo->max_over_ranks(t, comm.rank) : t; };
            Vector solution, rhs;
            system_matrix.initialize_dof_vector(solution);
            rhs.reinit(solution);
            system_matrix.compute_rhs(rhs);
            cudaDeviceSynchronize();
            std::printf("Total setup time: %g\n\n", since(t0));

            Row row{mesh.n_global_active_cells(), mesh.n_dofs()};
            cudaStream_t solve_stream;  // blocking stream: ordered with the default-stream work, and capturable
            check_cuda(cudaStreamCreate(&solve_stream), "cudaStreamCreate");
            double time_cg = 1e10;
            for (unsigned i = 0; i < 10; ++i) {
                ReductionControl solver_control(1000000000, 1e-16, 1e-9);
                SolverCG cg(solver_control);
                cg.set_stream(solve_stream);
                solution = 0;
                cudaDeviceSynchronize();
                auto t = clk::now();
                cg.solve(system_matrix, solution, rhs, PreconditionIdentity());
                cudaDeviceSynchronize();
                const double dt = max_over_ranks(since(t));
                time_cg = std::min(time_cg, dt);
                row.its = solver_control.last_step();
                row.red = std::pow(solver_control.last_value() / solver_control.initial_value(), 1. / solver_control.last_step());
                std::printf("Time solve CG              %g\n", dt);
            }
            cudaStreamDestroy(solve_stream);
            const unsigned n_mv = mesh.n_dofs() < 10000000 ? 200 : 50;
            auto best_of = [&](bool ghost_on, bool comp_on, bool plain) {
                double best = 1e10;
                for (unsigned i = 0; i < 5; ++i) {
                    cudaDeviceSynchronize();
                    auto t = clk::now();
                    for (unsigned k = 0; k < n_mv; ++k) {
                        if (plain) system_matrix.vmult(solution, rhs);
                        else system_matrix.vmult_dummy(solution, rhs, ghost_on, comp_on);
                    }
                    cudaDeviceSynchronize();
                    best = std::min(best, max_over_ranks(since(t)) / n_mv);
                }
                return best;
            };
            row.mv = best_of(true, true, true);
            std::printf("Best timings for ndof = %llu   mv %g    CG total %g   CG per iter. %g   [%.3f GDoF/s apply, %.3f GDoF/s CG]\n",
                        row.dofs, row.mv, time_cg, time_cg / row.its, 1e-9 * row.dofs / row.mv, 1e-9 * row.dofs * row.its / time_cg);
            row.cg_tot = time_cg; row.cg_per_it = time_cg / row.its;
            row.both = best_of(true, true, false); row.ghost = best_of(true, false, false); row.comp = best_of(false, true, false);
            table.push_back(row);
            std::printf("\n cells    dofs    matvec   CG_tot_time CG_time/iters cg_its cg_reduction \n");
            for (const Row &r : table)
                std::printf("%7llu %9llu %.3e   %.3e     %.3e %6u    %.3e \n", r.cells, r.dofs, r.mv, r.cg_tot, r.cg_per_it, r.its, r.red);
            std::printf("\n cells    dofs   mv_ghost_and_compute mv_compute_only mv_ghost_only \n");
            for (const Row &r : table) std::printf("%7llu %9llu           %.4e      %.4e    %.4e \n", r.cells, r.dofs, r.both, r.comp, r.ghost);
            std::printf("\n");
        }
    }
};

template <int p>
void run_degree(int nq_offset, std::size_t mn, std::size_t mx, Quadrature quad)
{
    if (nq_offset == 2) LaplaceProblem<p, p + 2>().run(mn, mx, quad);
    else LaplaceProblem<p, p + 1>().run(mn, mx, quad);
}

int main(int argc, char **argv)
{
    try {
        unsigned degree = 1;
        std::size_t minsize = 1, maxsize = static_cast<std::size_t>(-1);
        int nq_offset = 2;
        Quadrature quad = Quadrature::Gauss;
        if (argc == 1) { std::cout << "Expected at least one argument.\nUsage:\n./bp3 degree minsize maxsize [nq_offset] [gauss|gll]\n"; return 1; }
        degree = std::atoi(argv[1]);
        if (argc > 2) minsize = std::atoll(argv[2]);
        if (argc > 3) maxsize = std::atoll(argv[3]);
        if (argc > 4) nq_offset = std::atoi(argv[4]);
        if (argc > 5 && std::string(argv[5]) == "gll") { quad = Quadrature::GaussLobatto; nq_offset = 1; }
        switch (degree) {  // LaplaceRunTime<dim, 1, 8> (bp3.cc:536-557)
            case 1: run_degree<1>(nq_offset, minsize, maxsize, quad); break;
            case 2: run_degree<2>(nq_offset, minsize, maxsize, quad); break;
            case 3: run_degree<3>(nq_offset, minsize, maxsize, quad); break;
            case 4: run_degree<4>(nq_offset, minsize, maxsize, quad); break;
            case 5: run_degree<5>(nq_offset, minsize, maxsize, quad); break;
            case 6: run_degree<6>(nq_offset, minsize, maxsize, quad); break;
            case 7: run_degree<7>(nq_offset, minsize, maxsize, quad); break;
            case 8: run_degree<8>(nq_offset, minsize, maxsize, quad); break;
            default: std::cerr << "degree outside 1..8: no work" << std::endl;
        }
    } catch (std::exception &exc) {  // bp3.cc:600-624
        std::cerr << "\n\n----------------------------------------------------\nException on processing: \n"
                  << exc.what() << "\nAborting!\n----------------------------------------------------" << std::endl;
        return 1;
    }
    return 0;
}
