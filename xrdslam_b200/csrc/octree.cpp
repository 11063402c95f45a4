// Sparse voxel octree for Vox-Fusion: host-side map structure with the reference's node
// numbering.
//
// Replaces third_party/sparse_octree (reference @ f0366f20): Octree::init / insert
// (src/octree.cpp:35-115), get_centres_and_children (:297-346), find_octant (:155-175),
// Morton encode/decode (include/utils.h:64-110).  The node id -- which is ALSO the row of the
// embedding table a voxel corner uses (slam/models/sparse_voxel.py:309-316) -- is
// the creation order: points in input order x 8 corner offsets x root-to-leaf descent.  This
// implementation keeps nodes in one flat array indexed by that id (no pointers, no std::set).
// Parity: bit-exact against the reference's own svo.Octree built from its sources
// (oracle/_ref/svo.so), tests/test_voxfusion_cpu.py.
#include <stdint.h>
#include <string.h>

#include <deque>
#include <new>
#include <vector>

#include "xrdslam_b200.h"

namespace {

constexpr int kMaxBits = 21;
enum NodeType : int { NONLEAF = -1, SURFACE = 0, FEATURE = 1 };
const int kIncrX[8] = {0, 0, 0, 0, 1, 1, 1, 1};
const int kIncrY[8] = {0, 0, 1, 1, 0, 0, 1, 1};
const int kIncrZ[8] = {0, 1, 0, 1, 0, 1, 0, 1};

inline uint64_t level_mask(int i) {  // keeps the top 3*(i+1) of the 63 Morton bits
  uint64_t m = 0;
  for (int k = 0; k <= i; ++k) m |= 0x7000000000000000ull >> (3 * k);
  return m;
}
inline uint64_t spread3(uint64_t v) {
  uint64_t x = v & 0x1fffff;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}
inline uint64_t squeeze3(uint64_t v) {
  uint64_t x = v & 0x1249249249249249ull;
  x = (x | x >> 2) & 0x10c30c30c30c30c3ull;
  x = (x | x >> 4) & 0x100f00f00f00f00full;
  x = (x | x >> 8) & 0x1f0000ff0000ffull;
  x = (x | x >> 16) & 0x1f00000000ffffull;
  x = (x | x >> 32) & 0x1fffff;
  return x;
}
inline uint64_t morton(int x, int y, int z) {
  return (spread3((uint64_t)(int64_t)x) | (spread3((uint64_t)(int64_t)y) << 1) |
          (spread3((uint64_t)(int64_t)z) << 2)) & level_mask(kMaxBits - 1);
}

struct Node {
  uint64_t code;
  uint32_t side;
  int type;
  int child[8];
};

}  // namespace

struct XrdOctree {
  int size;
  int max_level;
  std::vector<Node> nodes;

  int new_node() {
    Node n;
    n.code = 0; n.side = 0; n.type = NONLEAF;
    for (int i = 0; i < 8; ++i) n.child[i] = -1;
    nodes.push_back(n);
    return (int)nodes.size() - 1;
  }
  int find_leaf(int x, int y, int z) const {
    int n = 0;
    unsigned edge = size / 2;
    for (int d = 1; d <= max_level; edge /= 2, ++d) {
      const int cid = ((x & edge) > 0) + 2 * ((y & edge) > 0) + 4 * ((z & edge) > 0);
      const int c = nodes[n].child[cid];
      if (c < 0) return -1;
      n = c;
    }
    return n;
  }
};

extern "C" XrdOctree* xrd_octree_create(int grid_dim) {
  if (grid_dim < 2 || (grid_dim & (grid_dim - 1))) return nullptr;
  XrdOctree* t = new (std::nothrow) XrdOctree();
  if (!t) return nullptr;
  t->size = grid_dim;
  t->max_level = 0;
  while ((1 << t->max_level) < grid_dim) ++t->max_level;
  const int root = t->new_node();
  t->nodes[root].side = grid_dim;
  return t;
}

extern "C" void xrd_octree_destroy(XrdOctree* t) { delete t; }

extern "C" int xrd_octree_num_nodes(const XrdOctree* t) { return t ? (int)t->nodes.size() : XRD_E_NULL; }

extern "C" int xrd_octree_insert(XrdOctree* t, const int32_t* voxels, int n) {
  if (!t || (!voxels && n > 0)) return XRD_E_NULL;
  const unsigned shift = kMaxBits - t->max_level - 1;
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < 8; ++j) {
      const int x = voxels[3 * i] + kIncrX[j], y = voxels[3 * i + 1] + kIncrY[j],
                z = voxels[3 * i + 2] + kIncrZ[j];
      const uint64_t key = morton(x, y, z);
      int cur = 0;
      unsigned edge = t->size / 2;
      for (int d = 1; d <= t->max_level; edge /= 2, ++d) {
        const int cid = ((x & edge) > 0) + 2 * ((y & edge) > 0) + 4 * ((z & edge) > 0);
        int c = t->nodes[cur].child[cid];
        if (c < 0) {
          c = t->new_node();
          Node& nn = t->nodes[c];
          nn.code = key & level_mask(d + shift);
          nn.side = edge;
          const bool leaf = (d == t->max_level);
          nn.type = leaf ? (j == 0 ? SURFACE : FEATURE) : NONLEAF;
          t->nodes[cur].child[cid] = c;
        } else if (t->nodes[c].type == FEATURE && j == 0) {
          t->nodes[c].type = SURFACE;
        }
        cur = c;
      }
    }
  }
  return (int)t->nodes.size();
}

// voxels [N][4] (x, y, z, side) f32; children [N][8] f32 (node id or -1, FEATURE leaves
// excluded); features [N][8] i32 (ids of the 8 corner leaves of every SURFACE leaf, else -1).
// Rows of nodes the breadth-first walk does not reach (FEATURE leaves) stay (0,0,0,0)/-1/-1,
// exactly as in the reference.
extern "C" int xrd_octree_export(const XrdOctree* t, float* voxels, float* children,
                                 int32_t* features) {
  if (!t || !voxels || !children || !features) return XRD_E_NULL;
  const int N = (int)t->nodes.size();
  for (int i = 0; i < N * 4; ++i) voxels[i] = 0.f;
  for (int i = 0; i < N * 8; ++i) { children[i] = -1.f; features[i] = -1; }
  std::deque<int> queue;
  queue.push_back(0);
  while (!queue.empty()) {
    const int id = queue.front();
    queue.pop_front();
    const Node& nd = t->nodes[id];
    const int cx = (int)squeeze3(nd.code), cy = (int)squeeze3(nd.code >> 1),
              cz = (int)squeeze3(nd.code >> 2);
    voxels[id * 4 + 0] = (float)cx; voxels[id * 4 + 1] = (float)cy;
    voxels[id * 4 + 2] = (float)cz; voxels[id * 4 + 3] = (float)nd.side;
    if (nd.type == SURFACE)
      for (int i = 0; i < 8; ++i) {
        const int leaf = t->find_leaf(cx + kIncrX[i], cy + kIncrY[i], cz + kIncrZ[i]);
        if (leaf >= 0) features[id * 8 + i] = leaf;
      }
    for (int i = 0; i < 8; ++i) {
      const int c = nd.child[i];
      if (c >= 0 && t->nodes[c].type != FEATURE) {
        queue.push_back(c);
        children[id * 8 + i] = (float)c;
      }
    }
  }
  return N;
}
