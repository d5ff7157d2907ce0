// mesh_common.h -- helpers shared by the host-side mesh builders (mesh.cc, hangmesh.cc).
#pragma once
#include <cstdint>
#include <vector>

struct b200fe_boxmesh;

namespace b200fe {

// mesh.cc: owned local indices every other rank ghosts (ascending), from this rank's own view (minimal ghost sets)
int boxmesh_send_lists(const b200fe_boxmesh *mesh, std::vector<std::vector<uint32_t>> &send);

// mesh.cc: what the device-side expansion of the index table needs (operator.cu: b200fe_boxmesh_dof_indices_device) --
// per own cell the local index of the first DoF of each of its 27 entities and the cell coordinates, per lexicographic
// local DoF its (entity, index in entity); all pointers into the mesh object
struct BoxMeshTables {
    int p, dirichlet;
    int64_t n_cells_local, cells[3];
    const uint32_t *lbase;   // [n_cells_local][27]
    const int32_t *cell_xyz; // [n_cells_local][3]
    const int *l_ent, *l_idx;  // [(p+1)^3]
};
int boxmesh_tables(const b200fe_boxmesh *mesh, BoxMeshTables *t);

namespace meshdetail {

// OpenMP team of the host-side mesh builders.  Launchers export OMP_NUM_THREADS=1 to every rank (torchrun does), which
// would serialise a setup that is embarrassingly parallel: take cores_available / ranks_on_this_node instead
// (B200FE_SETUP_THREADS overrides).  mesh.cc.
void use_setup_threads();

// hierarchical entity order of a hex: 8 vertices, 12 lines, 6 quads, 1 interior.
// code per axis: 0 = low plane, 1 = interior, 2 = high plane  (x, y, z)
struct EntityTable {
    int code[27][3];
    int dim[27];  // number of interior axes
    int id_of_code[3][3][3];
    EntityTable()
    {
        int n = 0;
        auto add = [&](int x, int y, int z) {
            code[n][0] = x; code[n][1] = y; code[n][2] = z;
            dim[n] = (x == 1) + (y == 1) + (z == 1);
            id_of_code[x][y][z] = n++;
        };
        for (int v = 0; v < 8; ++v) add((v & 1) ? 2 : 0, (v & 2) ? 2 : 0, (v & 4) ? 2 : 0);
        for (int z = 0; z <= 2; z += 2) {  // lines 0-3 (z low), 4-7 (z high)
            add(0, 1, z); add(2, 1, z); add(1, 0, z); add(1, 2, z);
        }
        add(0, 0, 1); add(2, 0, 1); add(0, 2, 1); add(2, 2, 1);  // lines 8-11
        add(0, 1, 1); add(2, 1, 1); add(1, 0, 1); add(1, 2, 1); add(1, 1, 0); add(1, 1, 2);  // quads
        add(1, 1, 1);
    }
};
inline const EntityTable kEnt;

inline uint64_t morton3(uint32_t x, uint32_t y, uint32_t z, int nbits)
{
    uint64_t c = 0;
    for (int b = 0; b < nbits; ++b)
        c |= (uint64_t)((x >> b) & 1) << (3 * b) | (uint64_t)((y >> b) & 1) << (3 * b + 1) |
             (uint64_t)((z >> b) & 1) << (3 * b + 2);
    return c;
}


// lexicographic local DoF (a + nm*(b + nm*c), a <-> x) of FE_Q<3>(p) -> hierarchical entity and index inside it
// (lines run along their axis, x-faces (y fastest, z), y-faces (z fastest, x), z-faces (x fastest, y))
inline void lexicographic_entities(int p, std::vector<int> &l_ent, std::vector<int> &l_idx)
{
    const int nm = p + 1, m = p - 1;
    l_ent.resize(nm * nm * nm);
    l_idx.resize(nm * nm * nm);
    for (int c = 0; c < nm; ++c)
        for (int b = 0; b < nm; ++b)
            for (int a = 0; a < nm; ++a) {
                const int t[3] = {a == 0 ? 0 : a == p ? 2 : 1, b == 0 ? 0 : b == p ? 2 : 1, c == 0 ? 0 : c == p ? 2 : 1};
                const int e = kEnt.id_of_code[t[0]][t[1]][t[2]];
                int idx = 0;
                const int ia = a - 1, ib = b - 1, ic = c - 1;
                switch (kEnt.dim[e]) {
                    case 0: idx = 0; break;
                    case 1: idx = t[0] == 1 ? ia : t[1] == 1 ? ib : ic; break;
                    case 2:
                        if (t[0] != 1) idx = ib + m * ic;        // x-face: (y fastest, z)
                        else if (t[1] != 1) idx = ic + m * ia;   // y-face: (z fastest, x)
                        else idx = ia + m * ib;                  // z-face: (x fastest, y)
                        break;
                    default: idx = ia + m * (ib + m * ic);
                }
                const int l = a + nm * (b + nm * c);
                l_ent[l] = e; l_idx[l] = idx;
            }
}

// number of DoFs of entity e (0 for lines / quads / interior at p = 1)
inline int entity_size(int p, int e)
{
    const int m = p - 1;
    return kEnt.dim[e] == 0 ? 1 : kEnt.dim[e] == 1 ? m : kEnt.dim[e] == 2 ? m * m : m * m * m;
}

}  // namespace meshdetail
}  // namespace b200fe
