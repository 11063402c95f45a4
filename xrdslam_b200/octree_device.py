"""Sparse voxel octree of Vox-Fusion built ON THE DEVICE with the reference's node numbering
(third_party/sparse_octree/src/octree.cpp:51-115 insert, :297-346
get_centres_and_children).

The reference inserts voxels one by one on the host: for every point (input order), for each
of its 8 corner voxels, descend root -> leaf and create the missing nodes; a node's id -- which
is also the row of the embedding table -- is its creation rank.  The same ids fall
out of a data-parallel formulation:

  * a node is identified by its PATH (depth d, the d child indices = the top d Morton triples of
    the voxel coordinate's low `max_level` bits);
  * its creation time is (first insertion that visits it, depth), insertions ordered as
    (first occurrence of the voxel in the input, corner j);
  * ids of new nodes = number of existing nodes + rank of (first insertion, depth).

Everything is tensor arithmetic on the points' device (torch.unique / sort / searchsorted /
scatter-min are the only primitives; no host round trip, no per-point loop).  Bit-exact against
the host C++ octree (csrc/octree.cpp, itself bit-exact against the reference's compiled
svo.Octree): tests/test_voxfusion_cpu.py."""
from __future__ import annotations

import torch

MAX_BITS = 21
NONLEAF, SURFACE, FEATURE = -1, 0, 1
INCR = torch.tensor([[0, 0, 0], [0, 0, 1], [0, 1, 0], [0, 1, 1],
                     [1, 0, 0], [1, 0, 1], [1, 1, 0], [1, 1, 1]], dtype=torch.int64)


def _spread3(v):
    x = v & 0x1fffff
    x = (x | (x << 32)) & 0x1f00000000ffff
    x = (x | (x << 16)) & 0x1f0000ff0000ff
    x = (x | (x << 8)) & 0x100f00f00f00f00f
    x = (x | (x << 4)) & 0x10c30c30c30c30c3
    x = (x | (x << 2)) & 0x1249249249249249
    return x


def _squeeze3(v):
    x = v & 0x1249249249249249
    x = (x | (x >> 2)) & 0x10c30c30c30c30c3
    x = (x | (x >> 4)) & 0x100f00f00f00f00f
    x = (x | (x >> 8)) & 0x1f0000ff0000ff
    x = (x | (x >> 16)) & 0x1f00000000ffff
    x = (x | (x >> 32)) & 0x1fffff
    return x


def _morton(xyz):
    """63-bit Morton code of int64 coords [..., 3] (x lowest bit of every triple)."""
    return _spread3(xyz[..., 0]) | (_spread3(xyz[..., 1]) << 1) | (_spread3(xyz[..., 2]) << 2)


class DeviceOctree:
    def __init__(self, grid_dim=256, device='cpu'):
        assert grid_dim >= 2 and (grid_dim & (grid_dim - 1)) == 0
        self.size = grid_dim
        self.max_level = grid_dim.bit_length() - 1
        self.device = torch.device(device)
        i64 = dict(dtype=torch.int64, device=self.device)
        # node table, row = node id (creation order); row 0 = root
        self.code = torch.zeros(1, **i64)
        self.side = torch.full((1,), grid_dim, **i64)
        self.type = torch.full((1,), NONLEAF, **i64)
        self.child = torch.full((1, 8), -1, **i64)
        self.depth = torch.zeros(1, **i64)
        # path keys (depth << 24*... | prefix), sorted, with their node ids
        self.pk_sorted = torch.zeros(1, **i64)   # root: depth 0, prefix 0
        self.pk_ids = torch.zeros(1, **i64)

    def num_nodes(self):
        return int(self.code.shape[0])

    # path key of depth d for low-bit Morton codes m (3 * max_level bits)
    def _pk(self, m, d):
        return (d << (3 * self.max_level)) | (m >> (3 * (self.max_level - d)))

    def _lookup(self, pk):
        """node id of every path key, -1 where absent."""
        pos = torch.searchsorted(self.pk_sorted, pk).clamp_(max=self.pk_sorted.shape[0] - 1)
        hit = self.pk_sorted[pos] == pk
        return torch.where(hit, self.pk_ids[pos], torch.full_like(pos, -1))

    def insert(self, voxels):
        """voxels: [n,3] integer voxel coordinates (floor(p / voxel_size)) in input order."""
        dev, L = self.device, self.max_level
        v = voxels.to(dev).to(torch.int64).reshape(-1, 3)
        n = v.shape[0]
        if n == 0:
            return self.num_nodes()
        # (a) distinct voxels in order of first appearance
        o = 1 << 20  # coordinates are 21-bit two's complement in the reference's Morton code
        key3 = ((v[:, 0] + o) << 42) | ((v[:, 1] + o) << 21) | (v[:, 2] + o)
        uniq, inv = torch.unique(key3, return_inverse=True)
        first = torch.full((uniq.shape[0],), n, dtype=torch.int64, device=dev)
        first.scatter_reduce_(0, inv, torch.arange(n, device=dev), 'amin')
        vu = v[torch.sort(first)[0]]                              # [U,3]
        U = vu.shape[0]
        # (b) insertions k' = (voxel rank, corner j); nodes visited at depth 1..L
        corners = vu[:, None, :] + INCR.to(dev)[None]             # [U,8,3]
        full = _morton(corners).reshape(-1)                       # [U*8] 63-bit codes
        low = _morton(corners & (self.size - 1)).reshape(-1)      # path bits only
        kk = torch.arange(U * 8, device=dev)
        d = torch.arange(1, L + 1, device=dev)
        pk = self._pk(low[:, None], d[None, :])                   # [U*8, L]
        # (c)+(d) new nodes: unseen path keys, creation time = (first k', depth)
        pk_f = pk.reshape(-1)
        k_f = kk[:, None].expand(-1, L).reshape(-1)
        d_f = d[None, :].expand(U * 8, -1).reshape(-1)
        new = self._lookup(pk_f) < 0
        pk_n, k_n, d_n = pk_f[new], k_f[new], d_f[new]
        if pk_n.numel():
            upk, inv2 = torch.unique(pk_n, return_inverse=True)
            fk = torch.full((upk.shape[0],), U * 8, dtype=torch.int64, device=dev)
            fk.scatter_reduce_(0, inv2, k_n, 'amin')
            ud = upk >> (3 * L)
            order = torch.argsort(fk * (L + 1) + ud)              # creation order
            upk, fk, ud = upk[order], fk[order], ud[order]
            N0 = self.num_nodes()
            ids = N0 + torch.arange(upk.shape[0], device=dev)
            shift = MAX_BITS - L - 1
            # code = creator's Morton key & level_mask(depth + shift): top 3*(d+shift+1) bits
            keep = 3 * (ud + shift + 1)
            mask = ((torch.ones_like(keep) << keep) - 1) << (63 - keep)
            code = full[fk] & mask
            side = self.size >> ud
            typ = torch.where(ud == L, torch.where(fk % 8 == 0, SURFACE, FEATURE),
                              torch.full_like(ud, NONLEAF))
            self.code = torch.cat([self.code, code])
            self.side = torch.cat([self.side, side])
            self.type = torch.cat([self.type, typ])
            self.depth = torch.cat([self.depth, ud])
            self.child = torch.cat([self.child, torch.full((upk.shape[0], 8), -1,
                                                           dtype=torch.int64, device=dev)])
            allk = torch.cat([self.pk_sorted, upk])
            alli = torch.cat([self.pk_ids, ids])
            o = torch.argsort(allk)
            self.pk_sorted, self.pk_ids = allk[o], alli[o]
            # (f) parent links
            prefix = upk & ((1 << (3 * L)) - 1)
            parent = self._lookup(((ud - 1) << (3 * L)) | (prefix >> 3))
            self.child[parent, prefix & 7] = ids
        # (e) a leaf reached by a voxel's own coordinate (corner 0) is a SURFACE leaf
        own = self._lookup(self._pk(low.reshape(U, 8)[:, 0], L))
        self.type[own] = SURFACE
        return self.num_nodes()

    def export(self):
        """-> voxels f32 [N,4] (x, y, z, side), children f32 [N,8], features i32 [N,8] with the
        semantics of get_centres_and_children: rows of FEATURE leaves stay 0 / -1 / -1."""
        dev, L = self.device, self.max_level
        N = self.num_nodes()
        live = self.type != FEATURE
        xyz = torch.stack([_squeeze3(self.code), _squeeze3(self.code >> 1),
                           _squeeze3(self.code >> 2)], -1)
        voxels = torch.cat([xyz, self.side[:, None]], -1).to(torch.float32)
        voxels = torch.where(live[:, None], voxels, torch.zeros_like(voxels))
        ch = self.child
        ch_type = self.type[ch.clamp(min=0)]
        ok = (ch >= 0) & (ch_type != FEATURE) & live[:, None]
        children = torch.where(ok, ch, torch.full_like(ch, -1)).to(torch.float32)
        features = torch.full((N, 8), -1, dtype=torch.int64, device=dev)
        surf = torch.nonzero(self.type == SURFACE).flatten()
        if surf.numel():
            c = xyz[surf][:, None, :] + INCR.to(dev)[None]        # [S,8,3]
            low = _morton(c & (self.size - 1))
            features[surf] = self._lookup(self._pk(low, L).reshape(-1)).reshape(-1, 8)
        return voxels, children, features.to(torch.int32)
