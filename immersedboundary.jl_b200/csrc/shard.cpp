// Rank-local shards of a Domain for one-process-per-GPU runs.
//
// The reference's own partition abstraction is a halo scheme: disjoint `image` sets (contiguous cell ranges,
// src/ImmersedBoundary.jl:594), a read-only 2-deep face-adjacency skirt (:610-619), results scattered to the
// image only (:857-859).  A rank owns a contiguous range of blocks (hence of cells); its local domain holds
// those blocks followed by every block that contains a skirt cell or an image-stencil donor of an owned ghost
// (Boundary.image_domain, :440-447).  Only the cells that are actually read are exchanged.
#include "ibx_internal.h"

#include <set>
#include <unordered_map>

namespace ibx {
ibx_domain* find_domain(const ibx_domain* d);
ibx_domain* register_domain(std::shared_ptr<ibx_domain> d);
void mesh_block_cells(const ibx_mesh& m, int64_t b, float* centers, float* widths);

// host mirror of the device-side neighbour lookup in fused.cu
static int host_neighbors(const ibx_domain& D, int64_t cpb, int64_t cell, int d, int side, int64_t* out) {
  int nd = D.nd, bs = D.block_size;
  int64_t b = cell / cpb;
  int l = (int)(cell - b * cpb);
  int ii[3] = {0, 0, 0};
  for (int k = 0; k < nd; ++k) { ii[k] = l % bs; l /= bs; }
  auto encode = [&](int64_t blk, const int* jj) {
    int64_t v = 0;
    for (int k = nd - 1; k >= 0; --k) v = v * bs + jj[k];
    return blk * cpb + v;
  };
  if ((side == 0 && ii[d] > 0) || (side == 1 && ii[d] < bs - 1)) {
    int64_t stride = 1;
    for (int k = 0; k < d; ++k) stride *= bs;
    out[0] = side ? cell + stride : cell - stride;
    return 1;
  }
  const BlockFace& bf = D.block_faces[(size_t)b * 2 * nd + 2 * d + side];
  int jj[3] = {ii[0], ii[1], ii[2]};
  jj[d] = side ? 0 : bs - 1;
  int t1 = d == 0 ? 1 : 0, t2 = nd == 3 ? (d == 2 ? 1 : 2) : t1;
  switch (bf.kind) {
    case 1:
      out[0] = encode(bf.nb[0], jj);
      return 1;
    case 2:
      jj[t1] = (ii[t1] + bf.sub[0] * bs) >> 1;
      if (nd == 3) jj[t2] = (ii[t2] + bf.sub[1] * bs) >> 1;
      out[0] = encode(bf.nb[0], jj);
      return 1;
    case 3: {
      int half = bs >> 1;
      int s1 = ii[t1] >= half, s2 = nd == 3 ? (ii[t2] >= half) : 0;
      int64_t nb = bf.nb[s1 + 2 * s2];
      int b1 = 2 * (ii[t1] - s1 * half), b2 = nd == 3 ? 2 * (ii[t2] - s2 * half) : 0;
      int cnt = nd == 3 ? 4 : 2;
      for (int q = 0; q < cnt; ++q) {
        jj[t1] = b1 + (q & 1);
        if (nd == 3) jj[t2] = b2 + (q >> 1);
        out[q] = encode(nb, jj);
      }
      return cnt;
    }
    default:
      return 0;  // domain box (or absent block): no other cell
  }
}

}  // namespace ibx

using namespace ibx;

extern "C" {

int ibx_domain_shard(const ibx_domain* gh, int rank, int nranks, ibx_domain** out) {
  IBX_TRY
  ibx_domain* Gp = find_domain(gh);
  IBX_REQUIRE(Gp != nullptr, "unknown domain handle");
  const ibx_domain& G = *Gp;
  IBX_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "rank out of range");
  IBX_REQUIRE(G.two_to_one, "sharding needs a mesh whose block contacts are same-level or 2:1");
  IBX_REQUIRE(G.ncells < (int64_t)2147483647, "more than 2^31-1 cells: the Int32 request / send lists would wrap");
  int nd = G.nd, bs = G.block_size;
  int64_t cpb = 1;
  for (int d = 0; d < nd; ++d) cpb *= bs;
  int64_t nb = G.ncells / cpb;
  IBX_REQUIRE(nb >= nranks, "fewer blocks than ranks");
  auto first_block = [&](int r) { return nb * r / nranks; };
  int64_t b0 = first_block(rank), b1 = first_block(rank + 1);
  int64_t c0 = b0 * cpb, c1 = b1 * cpb;
  auto owned = [&](int64_t c) { return c >= c0 && c < c1; };
  // --- skirt: two rounds of face adjacency from the owned range (src/ImmersedBoundary.jl:610-619)
  std::set<int64_t> need;  // global ids of non-owned cells that are read
  {
    std::vector<int64_t> frontier;
    // ring 1: neighbours of owned cells in blocks that have a non-owned contact
    for (int64_t b = b0; b < b1; ++b) {
      bool boundary_block = false;
      for (int f = 0; f < 2 * nd; ++f) {
        const BlockFace& bf = G.block_faces[(size_t)b * 2 * nd + f];
        int cnt = bf.kind == 3 ? (nd == 3 ? 4 : 2) : (bf.kind == 1 || bf.kind == 2 ? 1 : 0);
        for (int q = 0; q < cnt; ++q)
          if (bf.nb[q] < b0 || bf.nb[q] >= b1) boundary_block = true;
      }
      if (!boundary_block) continue;
      for (int64_t l = 0; l < cpb; ++l) {
        int64_t c = b * cpb + l;
        for (int d = 0; d < nd; ++d)
          for (int side = 0; side < 2; ++side) {
            int64_t nbr[4];
            int cnt = host_neighbors(G, cpb, c, d, side, nbr);
            for (int q = 0; q < cnt; ++q)
              if (!owned(nbr[q]) && need.insert(nbr[q]).second) frontier.push_back(nbr[q]);
          }
      }
    }
    // ring 2: neighbours of ring-1 cells
    for (int64_t c : frontier)
      for (int d = 0; d < nd; ++d)
        for (int side = 0; side < 2; ++side) {
          int64_t nbr[4];
          int cnt = host_neighbors(G, cpb, c, d, side, nbr);
          for (int q = 0; q < cnt; ++q)
            if (!owned(nbr[q])) need.insert(nbr[q]);
        }
  }
  // --- donors of owned ghosts
  for (const auto& F : G.boundaries)
    for (const auto& B : F.parts)
      for (size_t g = 0; g < B.ghost.size(); ++g) {
        if (!owned(B.ghost[g])) continue;
        for (int32_t k = B.ptr[g]; k < B.ptr[g + 1]; ++k) {
          int64_t c = B.image_domain[B.idx[k]];
          if (!owned(c)) need.insert(c);
        }
      }
  // --- local block list: owned blocks, then halo blocks by ascending global id
  std::set<int64_t> halo_blocks;
  for (int64_t c : need) halo_blocks.insert(c / cpb);
  // A halo block's face towards FINER blocks names 2^(nd-1) blocks, of which only those holding skirt cells are in
  // the set so far.  The sensor of a skirt cell next to such a face reads one 2^(nd-1) group, which lies in a single
  // (present) fine block; the siblings are added as placeholders so that the face entry stays addressable.
  {
    std::set<int64_t> extra;
    for (int64_t hb : halo_blocks)
      for (int f = 0; f < 2 * nd; ++f) {
        const BlockFace& bf = G.block_faces[(size_t)hb * 2 * nd + f];
        if (bf.kind != 3) continue;
        int cnt = nd == 3 ? 4 : 2;
        bool any = false;
        for (int q = 0; q < cnt; ++q) any |= (bf.nb[q] >= b0 && bf.nb[q] < b1) || halo_blocks.count(bf.nb[q]);
        if (any)
          for (int q = 0; q < cnt; ++q)
            if (!(bf.nb[q] >= b0 && bf.nb[q] < b1)) extra.insert(bf.nb[q]);
      }
    halo_blocks.insert(extra.begin(), extra.end());
  }
  std::vector<int64_t> lblocks;
  for (int64_t b = b0; b < b1; ++b) lblocks.push_back(b);
  for (int64_t b : halo_blocks) lblocks.push_back(b);
  std::unordered_map<int64_t, int32_t> g2l_block;
  for (size_t k = 0; k < lblocks.size(); ++k) g2l_block[lblocks[k]] = (int32_t)k;
  auto g2l_cell = [&](int64_t c) -> int64_t {
    auto it = g2l_block.find(c / cpb);
    return it == g2l_block.end() ? -1 : (int64_t)it->second * cpb + (c % cpb);
  };
  auto L = std::make_shared<ibx_domain>();
  L->nd = nd;
  L->block_size = bs;
  L->mesh = G.mesh;
  int64_t nlb = (int64_t)lblocks.size();
  L->ncells = nlb * cpb;
  L->two_to_one = true;
  L->centers.resize((size_t)L->ncells * nd);
  L->widths.resize((size_t)L->ncells * nd);
  L->block_faces.resize((size_t)nlb * 2 * nd);
  L->block_h.resize((size_t)nlb * nd);
  Shard& S = L->shard;
  S.active = true;
  S.rank = rank;
  S.nranks = nranks;
  S.owned_start = c0;
  S.n_owned = c1 - c0;
  S.n_halo = L->ncells - S.n_owned;
  S.local_to_global.resize((size_t)L->ncells);
  for (int64_t k = 0; k < nlb; ++k) {
    int64_t gb = lblocks[k];
    mesh_block_cells(*G.mesh, gb, L->centers.data() + k * cpb * nd, L->widths.data() + k * cpb * nd);
    for (int d = 0; d < nd; ++d) L->block_h[k * nd + d] = G.block_h[gb * nd + d];
    for (int64_t l = 0; l < cpb; ++l) S.local_to_global[k * cpb + l] = (int32_t)(gb * cpb + l);
    for (int f = 0; f < 2 * nd; ++f) {
      BlockFace bf = G.block_faces[(size_t)gb * 2 * nd + f];
      int cnt = bf.kind == 3 ? (nd == 3 ? 4 : 2) : (bf.kind == 1 || bf.kind == 2 ? 1 : 0);
      bool present = true;
      for (int q = 0; q < cnt; ++q) {
        auto it = g2l_block.find(bf.nb[q]);
        if (it == g2l_block.end()) present = false; else bf.nb[q] = it->second;
      }
      if (!present) {  // neighbour block not held locally: only possible for halo blocks, whose results are unused
        bf.kind = 0;
        bf.nb[0] = bf.nb[1] = bf.nb[2] = bf.nb[3] = -1;
      }
      L->block_faces[(size_t)k * 2 * nd + f] = bf;
    }
  }
  // --- boundaries restricted to owned ghosts, donors in local ids
  for (const auto& F : G.boundaries) {
    BoundaryFamily LF;
    LF.name = F.name;
    BoundaryT LB;
    LB.ptr.push_back(0);
    std::vector<int32_t> gidx;
    for (const auto& B : F.parts)
      for (size_t g = 0; g < B.ghost.size(); ++g) {
        if (!owned(B.ghost[g])) continue;
        LB.ghost.push_back((int32_t)g2l_cell(B.ghost[g]));
        for (int d = 0; d < nd; ++d) { LB.proj.push_back(B.proj[g * nd + d]); LB.normals.push_back(B.normals[g * nd + d]); }
        LB.image_dist.push_back(B.image_dist[g]);
        LB.ghost_dist.push_back(B.ghost_dist[g]);
        for (int32_t k = B.ptr[g]; k < B.ptr[g + 1]; ++k) {
          int64_t lc = g2l_cell(B.image_domain[B.idx[k]]);
          if (lc < 0) throw std::runtime_error("internal error: donor cell missing from the rank-local domain");
          gidx.push_back((int32_t)lc);
          LB.w.push_back(B.w[k]);
        }
        LB.ptr.push_back((int32_t)gidx.size());
      }
    LB.image_domain = gidx;
    std::sort(LB.image_domain.begin(), LB.image_domain.end());
    LB.image_domain.erase(std::unique(LB.image_domain.begin(), LB.image_domain.end()), LB.image_domain.end());
    LB.idx.resize(gidx.size());
    for (size_t q = 0; q < gidx.size(); ++q)
      LB.idx[q] = (int32_t)(std::lower_bound(LB.image_domain.begin(), LB.image_domain.end(), gidx[q]) - LB.image_domain.begin());
    if (!LB.ghost.empty()) LF.parts.push_back(std::move(LB));
    L->boundaries.push_back(std::move(LF));
  }
  // --- receive lists per owner rank (ascending global id); send lists are filled by ibx_shard_set_send
  S.send_local.assign(nranks, {});
  S.recv_local.assign(nranks, {});
  S.d_send.assign(nranks, nullptr);
  S.d_recv.assign(nranks, nullptr);
  S.d_sendbuf.assign(nranks, nullptr);
  S.d_recvbuf.assign(nranks, nullptr);
  S.buf_cap.assign(nranks, 0);
  for (int64_t c : need) {
    int64_t b = c / cpb;
    int owner = (int)std::min<int64_t>(nranks - 1, (b * nranks + nranks - 1) / nb);
    while (owner > 0 && first_block(owner) > b) --owner;
    while (owner < nranks - 1 && first_block(owner + 1) <= b) ++owner;
    S.recv_local[owner].push_back((int32_t)g2l_cell(c));
  }
  *out = register_domain(L);
  return IBX_OK;
  IBX_CATCH
}

int ibx_shard_info(const ibx_domain* lh, int64_t* n_owned, int64_t* n_halo, int64_t* owned_start) {
  IBX_TRY
  ibx_domain* L = find_domain(lh);
  IBX_REQUIRE(L && L->shard.active, "not a rank-local shard");
  *n_owned = L->shard.n_owned;
  *n_halo = L->shard.n_halo;
  *owned_start = L->shard.owned_start;
  return IBX_OK;
  IBX_CATCH
}

int ibx_shard_tables(const ibx_domain* lh, int32_t* local_to_global) {
  IBX_TRY
  ibx_domain* L = find_domain(lh);
  IBX_REQUIRE(L && L->shard.active, "not a rank-local shard");
  std::copy(L->shard.local_to_global.begin(), L->shard.local_to_global.end(), local_to_global);
  return IBX_OK;
  IBX_CATCH
}

int ibx_halo_sizes(const ibx_domain* lh, int nranks, int64_t* send_counts, int64_t* recv_counts) {
  IBX_TRY
  ibx_domain* L = find_domain(lh);
  IBX_REQUIRE(L && L->shard.active, "not a rank-local shard");
  IBX_REQUIRE(nranks == L->shard.nranks, "nranks mismatch");
  for (int p = 0; p < nranks; ++p) {
    send_counts[p] = (int64_t)L->shard.send_local[p].size();
    recv_counts[p] = (int64_t)L->shard.recv_local[p].size();
  }
  return IBX_OK;
  IBX_CATCH
}

int ibx_halo_lists(const ibx_domain* lh, int peer, int32_t* send_local, int32_t* recv_local) {
  IBX_TRY
  ibx_domain* L = find_domain(lh);
  IBX_REQUIRE(L && L->shard.active, "not a rank-local shard");
  IBX_REQUIRE(peer >= 0 && peer < L->shard.nranks, "peer out of range");
  if (send_local) std::copy(L->shard.send_local[peer].begin(), L->shard.send_local[peer].end(), send_local);
  if (recv_local) std::copy(L->shard.recv_local[peer].begin(), L->shard.recv_local[peer].end(), recv_local);
  return IBX_OK;
  IBX_CATCH
}

// The cells peer `peer` asked this rank for, as GLOBAL cell ids in the order the peer will unpack them.
int ibx_shard_set_send(ibx_domain* lh, int peer, int64_t n, const int32_t* global_ids) {
  IBX_TRY
  ibx_domain* L = find_domain(lh);
  IBX_REQUIRE(L && L->shard.active, "not a rank-local shard");
  IBX_REQUIRE(peer >= 0 && peer < L->shard.nranks, "peer out of range");
  IBX_REQUIRE(!L->uploaded, "send lists must be set before ibx_domain_upload");
  Shard& S = L->shard;
  S.send_local[peer].resize((size_t)n);
  for (int64_t k = 0; k < n; ++k) {
    int64_t c = global_ids[k];
    IBX_REQUIRE(c >= S.owned_start && c < S.owned_start + S.n_owned, "peer requested a cell this rank does not own");
    S.send_local[peer][k] = (int32_t)(c - S.owned_start);
  }
  return IBX_OK;
  IBX_CATCH
}

}  // extern "C"
