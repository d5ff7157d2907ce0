// b200fe/operator.hpp -- header-only C++ host layer over the C ABI (include/b200fe.h).
//
// Same class / method names as the reference so that its drivers port by changing only the types:
//   b200fe::LaplaceOperator<dim, fe_degree, nq, number>   <-> Portable::LaplaceOperator
//        (CEED_bp/include/portable_laplace_operator.h:17-96): vmult, Tvmult, vmult_dummy,
//        initialize_dof_vector, m, n; compute_diagonal / get_matrix_diagonal_inverse
//        (bp5_kokkos/benchmark.cc:157-168, 218-251)
//   b200fe::Vector            <-> LinearAlgebra::distributed::Vector<double, MemorySpace::Default>
//   b200fe::ReductionControl  <-> dealii::ReductionControl        (bp3.cc:268)
//   b200fe::SolverCG          <-> dealii::SolverCG                (bp3.cc:269-278)
//   b200fe::BoxMesh           <-> Triangulation + DoFHandler + AffineConstraints of the drivers
//        (GridGenerator::subdivided_hyper_rectangle + refine_global + distribute_dofs, bp3.cc:452-488)
//   b200fe::HangingBoxMesh    <-> the same after one local refinement (hanging-node constraints; config C5)
// Needs the CUDA runtime only for device allocations (cudaMalloc / cudaMemcpy).
#pragma once
#include <cuda_runtime.h>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <initializer_list>
#include <stdexcept>
#include <string>
#include <vector>

#include "../b200fe.h"

namespace b200fe {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
struct NoConvergence : Error {  // SolverControl::NoConvergence
    using Error::Error;
};

inline void check(int rc)
{
    if (rc != B200FE_OK) throw Error(rc, std::string("b200fe: ") + b200fe_last_error());
}
inline void check_cuda(cudaError_t e, const char *what)
{
    if (e != cudaSuccess) throw Error(B200FE_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

template <typename T>
class DeviceArray {
  public:
    DeviceArray() = default;
    explicit DeviceArray(size_t n) { resize(n); }
    DeviceArray(const DeviceArray &) = delete;
    DeviceArray &operator=(const DeviceArray &) = delete;
    DeviceArray(DeviceArray &&o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr; o.n_ = 0; }
    DeviceArray &operator=(DeviceArray &&o) noexcept
    {
        if (this != &o) { release(); p_ = o.p_; n_ = o.n_; o.p_ = nullptr; o.n_ = 0; }
        return *this;
    }
    ~DeviceArray() { release(); }
    void resize(size_t n)
    {
        release();
        n_ = n;
        if (n) check_cuda(cudaMalloc(&p_, n * sizeof(T)), "cudaMalloc");
    }
    void upload(const T *h, size_t n)
    {
        if (n_ != n) resize(n);
        if (n) check_cuda(cudaMemcpy(p_, h, n * sizeof(T), cudaMemcpyHostToDevice), "cudaMemcpy H2D");
    }
    void download(T *h) const
    {
        if (n_) check_cuda(cudaMemcpy(h, p_, n_ * sizeof(T), cudaMemcpyDeviceToHost), "cudaMemcpy D2H");
    }
    T *data() { return p_; }
    const T *data() const { return p_; }
    size_t size() const { return n_; }

  private:
    void release() { if (p_) cudaFree(p_); p_ = nullptr; n_ = 0; }
    T *p_ = nullptr;
    size_t n_ = 0;
};

// owned DoFs followed by ghosts, like LinearAlgebra::distributed::Vector; vector-valued problems (BP2/BP4/BP6)
// store n_components such blocks one after the other (component-blocked)
class Vector {
  public:
    Vector() = default;
    void reinit(size_t n_owned, size_t n_ghost, int n_components = 1)
    {
        n_owned_ = n_owned; n_ghost_ = n_ghost; n_comp_ = n_components;
        d_.resize((n_owned + n_ghost) * n_components);
        *this = 0.0;
    }
    void reinit(const Vector &o) { reinit(o.n_owned_, o.n_ghost_, o.n_comp_); }
    int n_components() const { return n_comp_; }
    Vector &operator=(double v)
    {
        if (v != 0.0) throw Error(B200FE_ERR_INVALID_ARG, "Vector = s only implemented for s = 0");
        if (d_.size()) check_cuda(cudaMemset(d_.data(), 0, d_.size() * sizeof(double)), "cudaMemset");
        return *this;
    }
    double *get_values() { return d_.data(); }
    const double *get_values() const { return d_.data(); }
    size_t locally_owned_size() const { return n_owned_; }
    size_t size_with_ghosts() const { return n_owned_ + n_ghost_; }
    std::vector<double> to_host() const  // owned entries (of every component)
    {
        std::vector<double> h(d_.size());
        d_.download(h.data());
        for (int c = 1; c < n_comp_; ++c) std::memmove(h.data() + c * n_owned_, h.data() + c * (n_owned_ + n_ghost_), n_owned_ * sizeof(double));
        h.resize(n_owned_ * n_comp_);
        return h;
    }
    void from_host(const std::vector<double> &h)
    {
        check_cuda(cudaMemcpy(d_.data(), h.data(), std::min(h.size(), d_.size()) * sizeof(double), cudaMemcpyHostToDevice), "cudaMemcpy H2D");
    }
    double l2_norm() const
    {
        DeviceArray<double> s(1);
        double sum = 0;
        for (int c = 0; c < n_comp_; ++c) {
            check(b200fe_sum_squares(n_owned_, d_.data() + c * (n_owned_ + n_ghost_), s.data(), nullptr));
            double h = 0;
            s.download(&h);
            sum += h;
        }
        return std::sqrt(sum);
    }

  private:
    DeviceArray<double> d_;
    size_t n_owned_ = 0, n_ghost_ = 0;
    int n_comp_ = 1;
};

class BoxMesh {
  public:
    BoxMesh(const int (&subdivisions)[3], int n_refine, int p, const double (&p1)[3], const double (&p2)[3],
            int n_ranks = 1, int rank = 0, int partition = B200FE_PARTITION_P4EST, int ghosts = B200FE_GHOSTS_MINIMAL,
            bool dirichlet = true)
    {
        b200fe_boxmesh_desc d{};
        for (int k = 0; k < 3; ++k) { d.subdivisions[k] = subdivisions[k]; d.p1[k] = p1[k]; d.p2[k] = p2[k]; }
        d.n_refine = n_refine; d.p = p; d.n_ranks = n_ranks; d.rank = rank; d.partition = partition; d.ghosts = ghosts;
        d.dirichlet = dirichlet ? 1 : 0;
        check(b200fe_boxmesh_create(&d, &h_));
        check(b200fe_boxmesh_info(h_, &info));
        degree = p;
        desc = d;
    }
    // mesh number `cycle` of the reference sweep (CEED_bp/src/bp3.cc:443-473)
    static BoxMesh bp3_cycle(unsigned cycle, int p, int n_ranks = 1, int rank = 0)
    {
        const int n_refine = cycle / 3, rem = cycle % 3;
        int sub[3];
        double p1[3] = {-1, -1, -1}, p2[3];
        for (int d = 0; d < 3; ++d) { sub[d] = d < rem ? 2 : 1; p2[d] = d < rem ? 2.8 : 0.9; }
        return BoxMesh(sub, n_refine, p, p1, p2, n_ranks, rank);
    }
    static unsigned long long bp3_projected_size(unsigned cycle, int p)
    {
        const unsigned n_refine = cycle / 3, rem = cycle % 3;
        unsigned long long s = 1;
        for (unsigned d = 0; d < 3; ++d) s *= (unsigned long long)((1u << n_refine) * (d < rem ? 2 : 1) * p + 1);
        return s;
    }
    BoxMesh(BoxMesh &&o) noexcept : info(o.info), desc(o.desc), degree(o.degree), h_(o.h_) { o.h_ = nullptr; }
    BoxMesh(const BoxMesh &) = delete;
    ~BoxMesh() { if (h_) b200fe_boxmesh_destroy(h_); }
    const b200fe_boxmesh *handle() const { return h_; }
    unsigned long long n_dofs() const { return info.n_dofs_global; }
    unsigned long long n_global_active_cells() const { return info.n_cells_global; }
    b200fe_exchange *make_exchange() const
    {
        b200fe_exchange *ex = nullptr;
        check(b200fe_exchange_create_box(&desc, &ex));
        return ex;
    }
    b200fe_boxmesh_info_t info{};
    b200fe_boxmesh_desc desc{};
    int degree = 0;

  private:
    b200fe_boxmesh *h_ = nullptr;
};

// BoxMesh whose cells [refine_lo, refine_hi) are refined once more: Triangulation with hanging nodes + DoFHandler +
// AffineConstraints (make_hanging_node_constraints + Dirichlet) of a deal.II driver (BASELINE config C5).
class HangingBoxMesh {
  public:
    HangingBoxMesh(const int (&subdivisions)[3], int n_refine, int p, const int (&refine_lo)[3], const int (&refine_hi)[3],
                   const double (&p1)[3], const double (&p2)[3], int n_ranks = 1, int rank = 0, bool dirichlet = true)
    {
        b200fe_hangmesh_desc d{};
        for (int k = 0; k < 3; ++k) {
            d.box.subdivisions[k] = subdivisions[k]; d.box.p1[k] = p1[k]; d.box.p2[k] = p2[k];
            d.refine_lo[k] = refine_lo[k]; d.refine_hi[k] = refine_hi[k];
        }
        d.box.n_refine = n_refine; d.box.p = p; d.box.n_ranks = n_ranks; d.box.rank = rank;
        d.box.partition = B200FE_PARTITION_P4EST; d.box.ghosts = B200FE_GHOSTS_MINIMAL; d.box.dirichlet = dirichlet ? 1 : 0;
        check(b200fe_hangmesh_create(&d, &h_));
        check(b200fe_hangmesh_info(h_, &info));
        degree = p;
        desc = d;
    }
    HangingBoxMesh(HangingBoxMesh &&o) noexcept : info(o.info), desc(o.desc), degree(o.degree), h_(o.h_) { o.h_ = nullptr; }
    HangingBoxMesh(const HangingBoxMesh &) = delete;
    ~HangingBoxMesh() { if (h_) b200fe_hangmesh_destroy(h_); }
    const b200fe_hangmesh *handle() const { return h_; }
    unsigned long long n_dofs() const { return info.n_dofs_global; }
    unsigned long long n_global_active_cells() const { return info.n_cells_global; }
    b200fe_exchange *make_exchange() const
    {
        b200fe_exchange *ex = nullptr;
        check(b200fe_exchange_create_hang(&desc, &ex));
        return ex;
    }
    b200fe_hangmesh_info_t info{};
    b200fe_hangmesh_desc desc{};
    int degree = 0;

  private:
    b200fe_hangmesh *h_ = nullptr;
};

// One process per GPU.  Rank / size come from the launcher's environment (RANK, WORLD_SIZE, LOCAL_RANK, MASTER_PORT as set
// by `python -m torch.distributed.run --no-python`; a plain run has none of them = one rank); the only thing that has to
// travel between the processes is the
// 128-byte NCCL id, published by rank 0 through a file (the reference uses MPI for everything, bp3.cc:564; there is no MPI
// in this image).
class Communicator {
  public:
    static Communicator from_environment()
    {
        Communicator c;
        process_start();
        c.rank = env_int({"RANK"}, 0);
        c.size = env_int({"WORLD_SIZE"}, 1);
        c.local_rank = env_int({"LOCAL_RANK"}, c.rank);
        if (c.size < 1 || c.rank < 0 || c.rank >= c.size) throw Error(B200FE_ERR_INVALID_ARG, "bad RANK / WORLD_SIZE in the environment");
        // per-launch name: the launcher's run id (torchrun exports TORCHELASTIC_RUN_ID to every rank) or, for other
        // launchers, B200FE_RUN_ID; plus the port and the user id.  A file left by an earlier or crashed launch then has a
        // different name or is replaced before use (rank 0 creates it with O_EXCL | O_NOFOLLOW after unlinking).
        const char *port = std::getenv("MASTER_PORT");
        const char *run = std::getenv("B200FE_RUN_ID");
        if (!run) run = std::getenv("TORCHELASTIC_RUN_ID");
        const char *dir = std::getenv("XDG_RUNTIME_DIR");
        c.id_file = std::string(dir && *dir ? dir : "/tmp") + "/b200fe_nccl_id_" + std::to_string((unsigned long)getuid()) + "_" +
                    (port ? port : "default") + "_" + (run ? run : "norun");
        for (char &ch : c.id_file)
            if (ch == ' ' || ch == ':') ch = '_';
        return c;
    }
    // collective: rank 0 creates the id and publishes it, the others wait for a file written during this launch
    std::string unique_id() const
    {
        std::string id(128, '\0');
        if (size == 1) return id;
        if (rank == 0) {
            std::remove(id_file.c_str());
            check(b200fe_comm_unique_id(&id[0]));
            const std::string tmp = id_file + ".tmp";
            std::remove(tmp.c_str());
            // never follow a planted symlink, never reuse an existing file
            const int fd = ::open(tmp.c_str(), O_WRONLY | O_CREAT | O_EXCL | O_NOFOLLOW, 0600);
            if (fd < 0 || ::write(fd, id.data(), 128) != 128) {
                if (fd >= 0) ::close(fd);
                throw Error(B200FE_ERR_COMM, "cannot write " + tmp);
            }
            ::close(fd);
            if (std::rename(tmp.c_str(), id_file.c_str()) != 0) throw Error(B200FE_ERR_COMM, "cannot publish " + id_file);
            return id;
        }
        const std::time_t start = process_start();
        for (int tries = 0; tries < 6000; ++tries) {  // up to 10 minutes
            struct stat st;
            // a regular file of ours, complete, and not older than this process (1 s of clock granularity)
            if (lstat(id_file.c_str(), &st) == 0 && S_ISREG(st.st_mode) && st.st_uid == getuid() && st.st_size == 128 &&
                st.st_mtime + 30 >= start) {  // written during this launch (ranks of one launch start within seconds)
                const int fd = ::open(id_file.c_str(), O_RDONLY | O_NOFOLLOW);
                if (fd >= 0) {
                    const ssize_t n = ::read(fd, &id[0], 128);
                    ::close(fd);
                    if (n == 128) return id;
                }
            }
            usleep(100000);
        }
        throw Error(B200FE_ERR_COMM, "timed out waiting for the NCCL id in " + id_file);
    }
    void cleanup() const { if (rank == 0 && size > 1) std::remove(id_file.c_str()); }
    int rank = 0, size = 1, local_rank = 0;
    std::string id_file;

  private:
    static std::time_t process_start()
    {
        static const std::time_t t = std::time(nullptr);  // first call = from_environment() at program start
        return t;
    }
    static int env_int(std::initializer_list<const char *> names, int fallback)
    {
        for (const char *n : names)
            if (const char *v = std::getenv(n)) return std::atoi(v);
        return fallback;
    }
};

// ghost exchange of one rank: peer tables from b200fe_exchange_* (no communication), NCCL communicator inside
class Halo {
  public:
    template <class Mesh>
    Halo(const Mesh &mesh, const Communicator &comm) : size_(comm.size)
    {
        b200fe_exchange *ex = mesh.make_exchange();
        int n_peers = 0;
        uint32_t n_send = 0, n_owned = 0, n_ghost = 0;
        b200fe_exchange_info(ex, &n_peers, &n_send, &n_owned, &n_ghost);
        std::vector<int32_t> peers(n_peers);
        std::vector<uint32_t> ro(n_peers), rc(n_peers), so(n_peers), sc(n_peers), si(n_send);
        b200fe_exchange_fill(ex, peers.data(), ro.data(), rc.data(), so.data(), sc.data(), si.data());
        b200fe_exchange_destroy(ex);
        const std::string id = comm.unique_id();
        b200fe_halo_desc d{};
        d.rank = comm.rank; d.n_ranks = comm.size; d.nccl_unique_id = id.data();
        d.n_owned = n_owned; d.n_ghost = n_ghost; d.n_peers = n_peers; d.peers = peers.data();
        d.recv_offset = ro.data(); d.recv_count = rc.data(); d.send_offset = so.data(); d.send_count = sc.data();
        d.h_send_indices = si.data(); d.n_send = n_send;
        check(b200fe_halo_create(&d, &h_));  // collective
        comm.cleanup();
    }
    Halo(const Halo &) = delete;
    ~Halo() { if (h_) b200fe_halo_destroy(h_); }
    b200fe_halo *handle() const { return h_; }
    // max over ranks of a host value (Utilities::MPI::min_max_avg(...).max of bp3.cc:301-302): every rank fills its slot
    double max_over_ranks(double v, int rank) const
    {
        DeviceArray<double> d(size_);
        std::vector<double> h(size_, 0.0);
        h[rank] = v;
        d.upload(h.data(), h.size());
        check(b200fe_halo_allreduce_sum(h_, d.data(), size_, nullptr));
        check_cuda(cudaDeviceSynchronize(), "allreduce");
        d.download(h.data());
        return *std::max_element(h.begin(), h.end());
    }

  private:
    b200fe_halo *h_ = nullptr;
    int size_ = 1;
};

// smooth mesh deformation of b200fe_boxmesh_nodes (deform_kind 1); amplitude 0 = the box itself
struct Deformation {
    double amplitude = 0.0, frequency = 0.0;
};

enum class Quadrature { Gauss, GaussLobatto };
// Geometric factors of the Laplace operators: Stored = G[cell][6][nq^3] streamed by the cell kernel (the reference's
// algorithm, portable_laplace_operator.h:239-302); OnTheFly (MappingQ1 meshes) = rebuilt inside the kernel
// from six constants per cell on undeformed (affine) meshes -- the separable "cartesian" kernel when every cell is an
// axis-aligned box and the operator is collocated -- or from the 8 vertices per cell on deformed (trilinear) ones.
enum class Geometry { Stored, OnTheFly };

template <int dim, int fe_degree, int nq, typename number = double>
class LaplaceOperator {
    static_assert(dim == 3, "the bake-off operators are three-dimensional");
    static_assert(std::is_same<number, double>::value, "FP64 only");

  public:
    // p_geo: degree of the mapping (MappingQ(p_geo)); op_kind: B200FE_OP_*
    LaplaceOperator(const BoxMesh &mesh, Quadrature quad = Quadrature::Gauss, int op_kind = B200FE_OP_LAPLACE, int p_geo = 1,
                    Deformation deform = {}, Geometry geometry = Geometry::Stored)
        : n_dofs_global_(mesh.n_dofs()), n_owned_(mesh.info.n_owned), n_ghost_(mesh.info.n_ghost), geometry_(geometry)
    {
        if (geometry == Geometry::OnTheFly && (p_geo != 1 || (op_kind != B200FE_OP_LAPLACE && deform.amplitude != 0.0)))
            throw Error(B200FE_ERR_INVALID_ARG, "Geometry::OnTheFly: MappingQ1 meshes; mass / Helmholtz operators on undeformed (axis-aligned) cells only");
        trilinear_ = geometry == Geometry::OnTheFly && deform.amplitude != 0.0;
        if (mesh.degree != fe_degree) throw Error(B200FE_ERR_INVALID_ARG, "mesh degree != fe_degree");
        const uint32_t nc = mesh.info.n_cells_local;
        std::vector<uint32_t> idx, con(mesh.info.n_constrained);  // idx stays empty: the table is expanded on the device
        check(b200fe_boxmesh_fill(mesh.handle(), nullptr, con.data(), nullptr, nullptr, nullptr, nullptr));
        idx_.resize((size_t)nc * nm3());
        if (nc) check(b200fe_boxmesh_dof_indices_device(mesh.handle(), idx_.data(), nullptr));
        DeviceArray<double> nodes((size_t)nc * 3 * ng3(p_geo));
        check(b200fe_boxmesh_nodes(mesh.handle(), p_geo, deform.amplitude != 0.0, deform.amplitude, deform.frequency, nodes.data(), nullptr));
        init(nc, idx, con, nodes, quad, op_kind, p_geo);
    }
    // two-level mesh with hanging nodes: the operator is C^T A C with identity on the constrained rows
    LaplaceOperator(const HangingBoxMesh &mesh, Quadrature quad = Quadrature::Gauss, int op_kind = B200FE_OP_LAPLACE, int p_geo = 1,
                    Deformation deform = {}, bool face_constraints = true)
        : n_dofs_global_(mesh.n_dofs()), n_owned_(mesh.info.n_owned), n_ghost_(mesh.info.n_ghost)
    {
        if (mesh.degree != fe_degree) throw Error(B200FE_ERR_INVALID_ARG, "mesh degree != fe_degree");
        const uint32_t nc = mesh.info.n_cells_local, nr = mesh.info.n_hanging_rows;
        std::vector<uint32_t> idx((size_t)nc * nm3()), con(mesh.info.n_constrained), hd(nr), hp(nr + 1), hc(mesh.info.n_hanging_entries);
        std::vector<double> hw(mesh.info.n_hanging_entries);
        check(b200fe_hangmesh_fill(mesh.handle(), idx.data(), con.data(), nullptr, nullptr, nullptr, nullptr, hd.data(), hp.data(), hc.data(), hw.data()));
        DeviceArray<double> nodes((size_t)nc * 3 * ng3(p_geo));
        check(b200fe_hangmesh_nodes(mesh.handle(), p_geo, deform.amplitude != 0.0, deform.amplitude, deform.frequency, nodes.data(), nullptr));
        init(nc, idx, con, nodes, quad, op_kind, p_geo);
        if (face_constraints && mesh.info.n_face_blocks) {  // tensor-product trace interpolation per coarse face (the fast path)
            const uint32_t nb = mesh.info.n_face_blocks;
            const size_t nm = fe_degree + 1, nf = 2 * fe_degree + 1;
            std::vector<uint32_t> fp((size_t)nb * nm * nm), fc((size_t)nb * nf * nf);
            std::vector<double> W(nf * nm);
            check(b200fe_hangmesh_fill_faces(mesh.handle(), fp.data(), fc.data()));
            check(b200fe_trace_weights(fe_degree, W.data()));
            check(b200fe_op_set_constraints(op_, nr, hd.data(), hp.data(), hc.data(), hw.data()));  // compute_diagonal reads the rows
            check(b200fe_op_set_face_constraints(op_, fe_degree, nb, fp.data(), fc.data(), W.data()));
        } else  // general AffineConstraints rows (CSR)
            check(b200fe_op_set_constraints(op_, nr, hd.data(), hp.data(), hc.data(), hw.data()));
    }
    LaplaceOperator(const LaplaceOperator &) = delete;
    ~LaplaceOperator() { if (op_) b200fe_op_destroy(op_); }

    void vmult(Vector &dst, const Vector &src) const  // vector-valued: the scalar operator on every component
    {
        if (src.n_components() == 1) check(b200fe_op_vmult(op_, dst.get_values(), src.get_values(), nullptr));
        else check(b200fe_op_vmult_components(op_, src.n_components(), dst.get_values(), src.get_values(), nullptr));
    }
    void Tvmult(Vector &dst, const Vector &src) const { vmult(dst, src); }
    void vmult_dummy(Vector &dst, const Vector &src, const bool ghost_exchange_on, const bool computation_on) const
    {
        check(b200fe_op_vmult_dummy(op_, dst.get_values(), src.get_values(), ghost_exchange_on, computation_on, nullptr));
    }
    void initialize_dof_vector(Vector &vec, int n_components = 1) const { vec.reinit(n_owned_, n_ghost_, n_components); }
    // AffineConstraints::distribute: hanging-node values of a solution from their parents
    void distribute(Vector &x) const
    {
        for (int c = 0; c < x.n_components(); ++c) check(b200fe_op_distribute(op_, x.get_values() + c * x.size_with_ghosts(), nullptr));
    }
    unsigned long long m() const { return n_dofs_global_; }
    unsigned long long n() const { return n_dofs_global_; }
    void compute_rhs(Vector &b) const  // bp3.cc:184-239 (f = 1 in every component)
    {
        for (int c = 0; c < b.n_components(); ++c) check(b200fe_op_rhs_one(op_, b.get_values() + c * b.size_with_ghosts(), nullptr));
    }
    void compute_diagonal()                                                                      // benchmark.cc:218-251
    {
        Vector d;
        initialize_dof_vector(d);
        check(b200fe_op_diagonal(op_, d.get_values(), nullptr));
        std::vector<double> h = d.to_host();
        for (double &x : h) x = x > 0 ? 1.0 / x : 1.0;
        inv_diag_.upload(h.data(), h.size());
    }
    const double *get_matrix_diagonal_inverse()
    {
        if (!inv_diag_.size()) compute_diagonal();
        return inv_diag_.data();
    }
    b200fe_op *handle() const { return op_; }
    // attach the ghost exchange (vmult then does update_ghost_values / compress(add) itself, and CG reduces its scalars)
    void set_halo(const Halo &halo) { check(b200fe_op_set_halo(op_, halo.handle())); }

  private:
    static constexpr size_t nm3() { return (size_t)(fe_degree + 1) * (fe_degree + 1) * (fe_degree + 1); }
    static size_t ng3(int p_geo) { return (size_t)(p_geo + 1) * (p_geo + 1) * (p_geo + 1); }
    void init(uint32_t nc, const std::vector<uint32_t> &idx, const std::vector<uint32_t> &con, const DeviceArray<double> &nodes,
              Quadrature quad, int op_kind, int p_geo)
    {
        const int nm = fe_degree + 1, qk = quad == Quadrature::Gauss ? B200FE_QUAD_GAUSS : B200FE_QUAD_GLL;
        const bool collocated = quad == Quadrature::GaussLobatto && nq == nm;
        std::vector<double> sv(nm * nq), cg(nq * nq);
        check(b200fe_basis_1d(fe_degree, nq, qk, sv.data(), cg.data(), nullptr, nullptr, nullptr));
        if (!idx.empty()) idx_.upload(idx.data(), idx.size());
        const size_t nq3 = (size_t)nq * nq * nq;
        const bool otf = geometry_ == Geometry::OnTheFly;
        if ((op_kind & B200FE_OP_LAPLACE) && !otf) G_.resize((size_t)nc * 6 * nq3);
        // JxW: streamed by the stored-geometry mass terms and read by compute_rhs; the separable kernels of undeformed meshes
        // need neither (det J is among the cell constants)
        if (!otf || trilinear_) JxW_.resize((size_t)nc * nq3);
        if (G_.size() || JxW_.size()) check(b200fe_geometry_from_nodes(p_geo, nq, qk, nc, nodes.data(), G_.data(), JxW_.data(), nullptr));
        std::vector<double> wts(nq), pts(nq);
        if (otf) {
            check(b200fe_basis_1d(fe_degree, nq, qk, nullptr, nullptr, nullptr, pts.data(), wts.data()));
            if (trilinear_) {  // the vertices ARE the geometry: keep a copy (the operator borrows it)
                cellX_.resize((size_t)nc * 24);
                check_cuda(cudaMemcpy(cellX_.data(), nodes.data(), (size_t)nc * 24 * sizeof(double), cudaMemcpyDeviceToDevice), "cudaMemcpy D2D");
            } else {
                cellG_.resize((size_t)nc * 8);
                check(b200fe_geometry_affine_from_nodes(nc, nodes.data(), cellG_.data(), nullptr));
            }
        }
        check_cuda(cudaDeviceSynchronize(), "geometry");
        b200fe_op_desc d{};
        d.p = fe_degree; d.nq = nq; d.op_kind = op_kind; d.collocated = collocated;
        d.n_cells = nc; d.n_owned = n_owned_; d.n_ghost = n_ghost_;
        d.h_shape_values = sv.data(); d.h_co_shape_gradients = cg.data();
        d.d_dof_indices = idx_.data(); d.d_G = G_.data(); d.d_JxW = JxW_.data();
        d.h_constrained = con.data(); d.n_constrained = (uint32_t)con.size();
        if (otf) {
            d.d_G = nullptr;
            d.h_weights = wts.data(); d.h_points = pts.data();
            if (trilinear_) d.d_cell_vertices = cellX_.data();
            else d.d_cell_G = cellG_.data();
        }
        check(b200fe_op_create(&d, &op_));
    }
    b200fe_op *op_ = nullptr;
    unsigned long long n_dofs_global_;
    uint32_t n_owned_, n_ghost_;
    DeviceArray<uint32_t> idx_;
    DeviceArray<double> G_, JxW_, inv_diag_, cellG_, cellX_;
    Geometry geometry_ = Geometry::Stored;
    bool trilinear_ = false;
};

class ReductionControl {
  public:
    ReductionControl(unsigned max_steps = 100, double tolerance = 1e-10, double reduction = 1e-2)
        : max_steps_(max_steps), tol_(tolerance), red_(reduction) {}
    unsigned last_step() const { return (unsigned)res_.iterations; }
    double last_value() const { return res_.final_residual; }
    double initial_value() const { return res_.initial_residual; }
    unsigned max_steps_;
    double tol_, red_;
    b200fe_cg_result res_{};
};

struct PreconditionIdentity {};

// dealii::SolverCG as the reference drivers call it.  One difference from deal.II: the solve always starts from x0 = 0 --
// the contents of x on entry are OVERWRITTEN, not used as an initial guess (dealii::SolverCG starts from r = b - A x).  The
// reference drivers zero the solution before every solve (bp3.cc:271, bp5_kokkos/benchmark.cc:364), so their iteration
// counts are the cold-start ones; for a warm start solve for the correction: A dx = b - A x0, x = x0 + dx.
class SolverCG {
  public:
    explicit SolverCG(ReductionControl &c, int check_every = 8) : control_(c), check_every_(check_every) {}
    // Run the solve on this (blocking) stream instead of the default stream; a non-default stream lets the
    // library replay CG iterations as a CUDA graph (small-problem latency).
    void set_stream(cudaStream_t s) { stream_ = s; }
    template <class Operator>
    void solve(const Operator &A, Vector &x, const Vector &b, const PreconditionIdentity &) { run(A.handle(), x, b, nullptr); }
    template <class Operator>
    void solve(const Operator &A, Vector &x, const Vector &b, const double *d_inverse_diagonal) { run(A.handle(), x, b, d_inverse_diagonal); }

  private:
    void run(b200fe_op *op, Vector &x, const Vector &b, const double *inv_diag)
    {
        const int max_it = control_.max_steps_ > 2000000000u ? 2000000000 : (int)control_.max_steps_;
        if (x.n_components() != b.n_components()) throw Error(B200FE_ERR_INVALID_ARG, "SolverCG: x and b differ in components");
        const int rc = b200fe_cg_solve_components(op, x.n_components(), x.get_values(), b.get_values(), inv_diag, control_.tol_,
                                                  control_.red_, max_it, check_every_, &control_.res_, stream_);
        if (rc == B200FE_ERR_NO_CONVERGENCE) throw NoConvergence(rc, b200fe_last_error());
        check(rc);
    }
    ReductionControl &control_;
    int check_every_;
    cudaStream_t stream_ = nullptr;
};

}  // namespace b200fe
