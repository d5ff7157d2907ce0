// comm_check.cc -- host-only check of the multi-rank plumbing of include/b200fe/operator.hpp (no GPU needed):
// every rank prints the NCCL id it obtained through b200fe::Communicator (rank 0 publishes, the others wait) and the sizes
// of its exchange lists for one box mesh and one hanging-node mesh.  tests/test_dist_cpu.py runs it as 2 and 3 processes.
//   RANK=r WORLD_SIZE=n MASTER_PORT=p ./comm_check <degree>
#include <b200fe/operator.hpp>

#include <cstdio>
#include <cstdlib>

using namespace b200fe;

template <class Mesh>
static void print_lists(const char *name, const Mesh &mesh)
{
    b200fe_exchange *ex = mesh.make_exchange();
    int n_peers = 0;
    uint32_t n_send = 0, n_owned = 0, n_ghost = 0;
    check(b200fe_exchange_info(ex, &n_peers, &n_send, &n_owned, &n_ghost));
    std::vector<int32_t> peers(n_peers);
    std::vector<uint32_t> ro(n_peers), rc(n_peers), so(n_peers), sc(n_peers), si(n_send);
    check(b200fe_exchange_fill(ex, peers.data(), ro.data(), rc.data(), so.data(), sc.data(), si.data()));
    b200fe_exchange_destroy(ex);
    unsigned long long sum = 0;
    for (uint32_t i : si) sum += i;
    std::printf("%s n_owned=%u n_ghost=%u n_peers=%d n_send=%u send_sum=%llu", name, n_owned, n_ghost, n_peers, n_send, sum);
    for (int k = 0; k < n_peers; ++k) std::printf(" peer%d:recv=%u,send=%u", peers[k], rc[k], sc[k]);
    std::printf("\n");
}

int main(int argc, char **argv)
{
    try {
        const int p = argc > 1 ? std::atoi(argv[1]) : 2;
        const Communicator comm = Communicator::from_environment();
        const std::string id = comm.unique_id();
        std::printf("rank=%d size=%d id=", comm.rank, comm.size);
        for (int i = 0; i < 16; ++i) std::printf("%02x", (unsigned char)id[i]);
        std::printf("\n");
        const int sub[3] = {2, 1, 1}, lo[3] = {1, 0, 0}, hi[3] = {3, 1, 2};
        const double p1[3] = {-1, -1, -1}, p2[3] = {2.8, 0.9, 0.9};
        BoxMesh box(sub, 1, p, p1, p2, comm.size, comm.rank);
        print_lists("box", box);
        HangingBoxMesh hang(sub, 1, p, lo, hi, p1, p2, comm.size, comm.rank);
        print_lists("hang", hang);
    } catch (std::exception &exc) {
        std::fprintf(stderr, "comm_check: %s\n", exc.what());
        return 1;
    }
    return 0;
}
