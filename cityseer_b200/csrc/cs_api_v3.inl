// Host side of the chain-contracted centrality_shortest kernel (cs_shortest3.cuh).  Included by cs_api.cu.
//
// Street networks are mostly chains: cityseer's decomposition cuts every street into ~20 m pieces, so nine nodes in ten
// have exactly two neighbours.  The upload therefore splits the node set into
//   * junctions  - everything that is not a plain pass-through node (degree != 2, one-way pieces, self-loops, ...), and
//   * interiors  - nodes with exactly two distinct neighbours a, b and one edge each way to both,
// and stores every maximal run of interiors between two junctions as one CHAIN: the per-segment travel seconds of both
// directions, contiguous in memory.  A junction-junction edge is a chain without interiors.  The search, the settle
// order, the predecessor rule and the dependency pass then run over junctions only (about a ninth of the reached nodes
// on the decomposed benchmark graph), and the interiors of a chain are produced by walking its seconds arrays with
// the same sequential f32 additions the reference performs node by node - bit-identical distances, no per-node queue,
// hash or sort traffic.  Node ids are renumbered: junctions first (Hilbert order), then interiors chain by chain, so
// the accumulator rows of a chain are contiguous.
//
// A graph qualifies when every non-loop directed edge has exactly one mutual twin (what io.network_structure_from_nx
// produces) and no junction has more than CS3_MAX_LINKS links; otherwise the arena kernel serves the call.

// position of (x, y) along a Hilbert curve over a 2^bits x 2^bits grid
static inline uint64_t hilbert_d(uint32_t x, uint32_t y, int bits) {
    uint64_t d = 0;
    for (uint32_t s = 1u << (bits - 1); s > 0; s >>= 1) {
        const uint32_t rx = (x & s) ? 1u : 0u, ry = (y & s) ? 1u : 0u;
        d += (uint64_t)s * s * ((3u * rx) ^ ry);
        if (ry == 0) {
            if (rx == 1) {
                x = s - 1 - x;
                y = s - 1 - y;
            }
            std::swap(x, y);
        }
    }
    return d;
}

// per-call flags by new id: dst[v] = src[orig_of_new[v]]
__global__ void cs_k_permute_u8(const uint8_t* src, const uint32_t* orig_of_new, uint8_t* dst, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[orig_of_new[i]];
}

static int build_v3_graph(cs_graph* g, uint32_t n, const uint8_t* node_exists, const double* xs, const double* ys,
                          const std::vector<uint32_t>& in_off, const std::vector<uint32_t>& out_off,
                          const std::vector<CsEdge>& in_rec, const std::vector<CsEdge>& out_rec,
                          const std::vector<float>& in_num, const std::vector<float>& out_num,
                          const std::vector<float>& in_imp, const std::vector<float>& weight) {
    g->v3_ok = false;
    const float MIN_NUM = 0.05f;  // shorter pieces are never contracted (f32 walks must stay strictly increasing; the kernel checks)
    // ---- every non-loop in-edge needs a mutual twin
    // in_rec[slot at v for u->v].meta: [7:0] position in u's out-list, [8] twin exists, [9] self-loop,
    // [21:16] 1 + position of the twin (v->u) inside u's in-list
    auto twin_slot = [&](uint32_t v, uint32_t j) -> int64_t {  // in-list slot (absolute) of the twin of in-edge j at v
        const CsEdge& r = in_rec[in_off[v] + j];
        const uint32_t back = (r.meta >> 16) & 0x3fu;
        if (!(r.meta & 0x100u) || back == 0) return -1;
        return (int64_t)in_off[r.nbr] + (back - 1);
    };
    for (uint32_t v = 0; v < n; ++v) {
        for (uint32_t j = 0; j < in_off[v + 1] - in_off[v]; ++j) {
            const CsEdge& r = in_rec[in_off[v] + j];
            if (r.meta & 0x200u) continue;
            const int64_t t = twin_slot(v, j);
            if (t < 0) return 0;
            const CsEdge& tr = in_rec[t];
            if (tr.nbr != v) return 0;
            const uint32_t tb = (tr.meta >> 16) & 0x3fu;
            if (!(tr.meta & 0x100u) || tb == 0 || tb - 1 != j) return 0;  // not mutual (parallel edges sharing a twin)
        }
    }
    // canonical flag of the directed edge stored at in-list slot s (absolute): look it up in the source's out-list
    auto canon_of_in = [&](uint32_t v, uint32_t s) -> uint32_t {
        const CsEdge& r = in_rec[s];
        (void)v;
        return (out_rec[out_off[r.nbr] + (r.meta & 0xffu)].meta >> 8) & 1u;
    };
    // ---- classification
    std::vector<uint8_t> interior(n, 0);
    for (uint32_t v = 0; v < n; ++v) {
        if (!node_exists[v]) continue;
        if (in_off[v + 1] - in_off[v] != 2 || out_off[v + 1] - out_off[v] != 2) continue;
        const CsEdge &i0 = in_rec[in_off[v]], &i1 = in_rec[in_off[v] + 1];
        const uint32_t a = i0.nbr, b = i1.nbr;
        if (a == b || a == v || b == v) continue;
        const CsEdge &o0 = out_rec[out_off[v]], &o1 = out_rec[out_off[v] + 1];
        if (!((o0.nbr == a && o1.nbr == b) || (o0.nbr == b && o1.nbr == a))) continue;
        bool good = true;
        for (uint32_t j = 0; j < 2 && good; ++j) {
            const float x = in_num[in_off[v] + j], y = out_num[out_off[v] + j];
            good = std::isfinite(x) && std::isfinite(y) && x >= MIN_NUM && y >= MIN_NUM;
            // exactly one canonical edge per piece (circuit rank counts each piece once)
            const int64_t t = twin_slot(v, j);
            good = good && t >= 0 && (canon_of_in(v, in_off[v] + j) + canon_of_in(in_rec[in_off[v] + j].nbr, (uint32_t)t) == 1u);
        }
        if (good) interior[v] = 1;
    }
    // ---- chains: walk from every junction through runs of interiors; runs longer than CS3_KMAX are cut by promoting
    //      an interior to a junction, and interior-only cycles are promoted entirely
    struct Chain {
        uint32_t A, B, k, first, last;
        std::vector<uint32_t> nodes;
    };
    std::vector<Chain> chains;
    std::vector<int32_t> chain_of(n, -1);
    {
        std::vector<uint32_t> todo;
        for (uint32_t v = 0; v < n; ++v)
            if (node_exists[v] && !interior[v]) todo.push_back(v);
        for (int round = 0; round < 2; ++round) {
            for (size_t w = 0; w < todo.size(); ++w) {
                const uint32_t v = todo[w];
                for (uint32_t j = 0; j < in_off[v + 1] - in_off[v]; ++j) {
                    const uint32_t x = in_rec[in_off[v] + j].nbr;
                    if (x == v || !interior[x] || chain_of[x] >= 0) continue;
                    Chain c;
                    c.A = v;
                    uint32_t prev = v, cur = x;
                    while (interior[cur]) {
                        if (c.nodes.size() == CS3_KMAX) {
                            interior[cur] = 0;  // cut the run here: cur becomes a junction
                            todo.push_back(cur);
                            break;
                        }
                        c.nodes.push_back(cur);
                        const uint32_t a = in_rec[in_off[cur]].nbr, b = in_rec[in_off[cur] + 1].nbr;
                        const uint32_t nxt = a == prev ? b : a;
                        prev = cur;
                        cur = nxt;
                    }
                    c.B = cur;
                    c.k = (uint32_t)c.nodes.size();
                    c.first = c.nodes.front();
                    c.last = c.nodes.back();
                    for (uint32_t nd : c.nodes) chain_of[nd] = (int32_t)chains.size();
                    chains.push_back(std::move(c));
                }
            }
            if (round == 0) {
                // interior-only cycles were never reached from a junction: promote all their nodes
                todo.clear();
                for (uint32_t v = 0; v < n; ++v)
                    if (interior[v] && chain_of[v] < 0) interior[v] = 0;
            }
        }
    }
    // ---- link lists of the junctions (in-list order, self-loops dropped)
    std::vector<uint32_t> junctions;
    for (uint32_t v = 0; v < n; ++v)
        if (node_exists[v] && !interior[v]) junctions.push_back(v);
    const uint32_t J = (uint32_t)junctions.size();
    if (J == 0) return 0;
    // link index of every in-list slot of a junction (0xff for self-loops)
    std::vector<uint8_t> link_idx(in_rec.size(), 0xff);
    std::vector<uint32_t> nlinks(J, 0), indeg(J, 0);
    for (uint32_t q = 0; q < J; ++q) {
        const uint32_t v = junctions[q];
        uint32_t c = 0;
        for (uint32_t j = 0; j < in_off[v + 1] - in_off[v]; ++j) {
            if (in_rec[in_off[v] + j].meta & 0x200u) {
                g->v3_loops = true;  // self-loops are visited edges of segment_centrality: served by the node-level kernel
                continue;
            }
            if (c >= CS3_MAX_LINKS) return 0;
            link_idx[in_off[v] + j] = (uint8_t)c++;
        }
        nlinks[q] = c;
        indeg[q] = in_off[v + 1] - in_off[v];
    }
    // direct junction-junction edges become chains without interiors (one per twin pair)
    std::vector<int32_t> slot_chain(in_rec.size(), -1);  // chain of every junction in-list slot
    std::vector<uint8_t> slot_dir(in_rec.size(), 0);
    struct Direct {
        uint32_t A, B, slotA, slotB, cnt;
    };
    std::vector<Direct> directs;
    for (uint32_t q = 0; q < J; ++q) {
        const uint32_t v = junctions[q];
        for (uint32_t j = 0; j < in_off[v + 1] - in_off[v]; ++j) {
            const uint32_t s = in_off[v] + j;
            const CsEdge& r = in_rec[s];
            if (r.meta & 0x200u) continue;
            if (interior[r.nbr]) {
                const int32_t c = chain_of[r.nbr];
                const Chain& ch = chains[c];
                // which end of the chain is this slot?  (a loop A == B with k >= 2 has distinct first / last)
                // which end of the chain is this slot?  (a loop A == B has k >= 2, hence distinct first / last nodes)
                slot_chain[s] = c;
                slot_dir[s] = (ch.A == v && ch.first == r.nbr) ? 0 : 1;
            } else if (slot_chain[s] < 0) {
                const int64_t t = twin_slot(v, j);
                Direct d;
                d.A = v;
                d.B = r.nbr;
                d.slotA = s;
                d.slotB = (uint32_t)t;
                d.cnt = canon_of_in(v, s) + canon_of_in(r.nbr, (uint32_t)t);
                slot_chain[s] = (int32_t)(chains.size() + directs.size());
                slot_dir[s] = 0;
                slot_chain[t] = slot_chain[s];
                slot_dir[t] = 1;
                directs.push_back(d);
            }
        }
    }
    // ---- renumbering: junctions along a Hilbert curve, interiors chain by chain (chains ordered by their first node)
    std::vector<uint64_t> hkey(n, ~0ull);
    {
        bool coords = xs != nullptr && ys != nullptr;
        double x0 = INFINITY, x1 = -INFINITY, y0 = INFINITY, y1 = -INFINITY;
        if (coords)
            for (uint32_t i = 0; i < n; ++i) {
                if (!node_exists[i]) continue;
                if (!std::isfinite(xs[i]) || !std::isfinite(ys[i])) {
                    coords = false;
                    break;
                }
                x0 = std::min(x0, xs[i]);
                x1 = std::max(x1, xs[i]);
                y0 = std::min(y0, ys[i]);
                y1 = std::max(y1, ys[i]);
            }
        const double span = coords ? std::max(std::max(x1 - x0, y1 - y0), 1e-9) : 1.0;
        for (uint32_t i = 0; i < n; ++i)
            if (node_exists[i])
                hkey[i] = coords ? hilbert_d((uint32_t)((xs[i] - x0) * 65535.0 / span), (uint32_t)((ys[i] - y0) * 65535.0 / span), 16)
                                 : (uint64_t)i;
    }
    std::vector<uint32_t> jorder(J);
    std::iota(jorder.begin(), jorder.end(), 0u);
    std::stable_sort(jorder.begin(), jorder.end(), [&](uint32_t a, uint32_t b) { return hkey[junctions[a]] < hkey[junctions[b]]; });
    std::vector<uint32_t> new_of_orig(n, 0xffffffffu), orig_of_new(n, 0);
    for (uint32_t q = 0; q < J; ++q) {
        new_of_orig[junctions[jorder[q]]] = q;
        orig_of_new[q] = junctions[jorder[q]];
    }
    const uint32_t NC = (uint32_t)chains.size();
    std::vector<uint32_t> corder(NC);
    std::iota(corder.begin(), corder.end(), 0u);
    std::stable_sort(corder.begin(), corder.end(), [&](uint32_t a, uint32_t b) { return hkey[chains[a].first] < hkey[chains[b].first]; });
    std::vector<uint32_t> ibase(NC, 0);
    uint32_t I = 0;
    for (uint32_t ci : corder) {
        ibase[ci] = I;
        for (uint32_t t = 0; t < chains[ci].k; ++t) {
            new_of_orig[chains[ci].nodes[t]] = J + I + t;
            orig_of_new[J + I + t] = chains[ci].nodes[t];
        }
        I += chains[ci].k;
    }
    {
        uint32_t next = J + I;
        for (uint32_t v = 0; v < n; ++v)
            if (new_of_orig[v] == 0xffffffffu) {
                new_of_orig[v] = next;
                orig_of_new[next++] = v;
            }
    }
    // ---- seconds numerators: per chain fwd[0..k] (wave A -> B, step t uses the edge m_{t+1} -> m_t), padding to a
    //      multiple of four floats, then bwdr[0..k] (wave B -> A, step t uses the edge m_{k-t} -> m_{k+1-t})
    const uint32_t C = NC + (uint32_t)directs.size();
    std::vector<uint32_t> soff(C, 0);
    std::vector<float> cnum;
    // segment_centrality reads, per visited piece, the length and impedance of the edge LEAVING the visiting node
    // (get_edge_length_unchecked(start = popped node, end = neighbour), centrality.rs:2234-2243): same block layout as the
    // seconds - entry t of the fwd array belongs to the piece the A -> B wave crosses in step t, visited from m_t
    std::vector<float> clen, cimp;
    auto in_slot_from = [&](uint32_t at, uint32_t from) -> uint32_t {  // in-list slot (absolute) of the edge from -> at
        for (uint32_t s2 = in_off[at]; s2 < in_off[at + 1]; ++s2)
            if (in_rec[s2].nbr == from) return s2;
        return in_off[at];
    };
    auto in_num_from = [&](uint32_t at, uint32_t from) -> float {  // numerator of the edge from -> at (at interior)
        return in_rec[in_off[at]].nbr == from ? in_num[in_off[at]] : in_num[in_off[at] + 1];
    };
    auto out_num_to = [&](uint32_t at, uint32_t to) -> float {  // numerator of the edge at -> to (at interior)
        return out_rec[out_off[at]].nbr == to ? out_num[out_off[at]] : out_num[out_off[at] + 1];
    };
    for (uint32_t c = 0; c < NC; ++c) {
        const Chain& ch = chains[c];
        const uint32_t k = ch.k;
        while (cnum.size() & 3u) cnum.push_back(0.f);  // blocks are fetched with 16-byte loads
        soff[c] = (uint32_t)cnum.size();
        auto node = [&](uint32_t t) { return t == 0 ? ch.A : t == k + 1 ? ch.B : ch.nodes[t - 1]; };
        std::vector<float> fwd(k + 1), bwd(k + 1);
        for (uint32_t t = 0; t <= k; ++t) {
            // edge m_{t+1} -> m_t and edge m_t -> m_{t+1}; read both at whichever end is an interior
            if (t < k) {  // m_{t+1} is an interior
                fwd[t] = out_num_to(node(t + 1), node(t));
                bwd[t] = in_num_from(node(t + 1), node(t));
            } else {  // m_k is an interior (k >= 1)
                fwd[t] = in_num_from(node(k), node(k + 1));
                bwd[t] = out_num_to(node(k), node(k + 1));
            }
        }
        clen.resize(cnum.size(), 0.f);
        cimp.resize(cnum.size(), 1.f);
        for (uint32_t t = 0; t <= k; ++t) {
            cnum.push_back(fwd[t]);
            // the wave from A crosses piece t from m_t: the edge m_t -> m_{t+1} is the twin of the in-edge m_{t+1} -> m_t
            const uint32_t s2 = in_slot_from(node(t), node(t + 1));
            clen.push_back(in_rec[s2].aux);
            cimp.push_back(in_imp[s2]);
        }
        while (cnum.size() & 3u) cnum.push_back(0.f);  // bwdr starts at cs3_pb(k): both arrays 16-byte aligned
        clen.resize(cnum.size(), 0.f);
        cimp.resize(cnum.size(), 1.f);
        for (uint32_t t = 0; t <= k; ++t) {
            cnum.push_back(bwd[k - t]);
            // the wave from B crosses piece k - t from m_{k+1-t}: the edge m_{k+1-t} -> m_{k-t}
            const uint32_t s2 = in_slot_from(node(k + 1 - t), node(k - t));
            clen.push_back(in_rec[s2].aux);
            cimp.push_back(in_imp[s2]);
        }
    }
    for (uint32_t d = 0; d < directs.size(); ++d) {
        while (cnum.size() & 3u) cnum.push_back(0.f);
        soff[NC + d] = (uint32_t)cnum.size();
        clen.resize(cnum.size(), 0.f);
        cimp.resize(cnum.size(), 1.f);
        cnum.push_back(in_num[directs[d].slotA]);  // B -> A: A's outward step
        clen.push_back(in_rec[directs[d].slotA].aux);
        cimp.push_back(in_imp[directs[d].slotA]);
        while (cnum.size() & 3u) cnum.push_back(0.f);
        clen.resize(cnum.size(), 0.f);
        cimp.resize(cnum.size(), 1.f);
        cnum.push_back(in_num[directs[d].slotB]);  // A -> B: B's outward step (at cs3_pb(0))
        clen.push_back(in_rec[directs[d].slotB].aux);
        cimp.push_back(in_imp[directs[d].slotB]);
    }
    while (cnum.size() & 3u) cnum.push_back(0.f);
    clen.resize(cnum.size(), 0.f);
    cimp.resize(cnum.size(), 1.f);
    // ---- link records by new junction id
    std::vector<uint32_t> jn_off(J + 1, 0);
    for (uint32_t q = 0; q < J; ++q) jn_off[q + 1] = jn_off[q] + nlinks[jorder[q]];
    std::vector<uint4> links(jn_off[J]);
    std::vector<uint2> jinfo(J);
    // per chain: link position at both ends (needed by the other end and by sources inside the chain)
    std::vector<uint32_t> posA(C, 0), posB(C, 0), endA(C, 0), endB(C, 0);
    for (uint32_t q = 0; q < J; ++q) {
        const uint32_t v = junctions[jorder[q]];
        for (uint32_t j = 0; j < in_off[v + 1] - in_off[v]; ++j) {
            const uint32_t s = in_off[v] + j;
            if (link_idx[s] == 0xff) continue;
            const int32_t c = slot_chain[s];
            if (slot_dir[s] == 0) {
                posA[c] = link_idx[s];
                endA[c] = q;
            } else {
                posB[c] = link_idx[s];
                endB[c] = q;
            }
        }
    }
    for (uint32_t q = 0; q < J; ++q) {
        const uint32_t v = junctions[jorder[q]];
        jinfo[q] = make_uint2(jn_off[q], nlinks[jorder[q]] | (indeg[jorder[q]] << 8));
        for (uint32_t j = 0; j < in_off[v + 1] - in_off[v]; ++j) {
            const uint32_t s = in_off[v] + j;
            if (link_idx[s] == 0xff) continue;
            const uint32_t c = (uint32_t)slot_chain[s];
            const uint32_t dir = slot_dir[s];
            const uint32_t k = c < NC ? chains[c].k : 0u;
            const uint32_t cnt = c < NC ? 1u : directs[c - NC].cnt;
            const uint32_t far = dir == 0 ? endB[c] : endA[c];
            const uint32_t paf = dir == 0 ? posB[c] : posA[c];
            links[jn_off[q] + link_idx[s]] =
                make_uint4(far, soff[c], c < NC ? ibase[c] : 0u, k | (dir << 4) | (paf << 5) | (cnt << 9));
        }
    }
    // ---- chain table for sources that are interiors: {soff, ibase, k, A} {B, posA, posB, -}
    std::vector<uint4> ctab(2 * (size_t)std::max<uint32_t>(NC, 1));
    std::vector<uint32_t> int_chain(std::max<uint32_t>(I, 1), 0);
    for (uint32_t c = 0; c < NC; ++c) {
        ctab[2 * c] = make_uint4(soff[c], ibase[c], chains[c].k, endA[c]);
        ctab[2 * c + 1] = make_uint4(endB[c], posA[c], posB[c], 0u);
        for (uint32_t t = 0; t < chains[c].k; ++t) int_chain[ibase[c] + t] = c;
    }
    std::vector<float> weight3(n);
    for (uint32_t v = 0; v < n; ++v) weight3[v] = weight[orig_of_new[v]];
    int rc = 0;
    rc |= upload(&g->d3_jinfo, jinfo);
    rc |= upload(&g->d3_links, links);
    rc |= upload(&g->d3_cnum, cnum);
    rc |= upload(&g->d3_csec, cnum);
    rc |= upload(&g->d3_clen, clen);
    rc |= upload(&g->d3_cimp, cimp);
    rc |= upload(&g->d3_ctab, ctab);
    rc |= upload(&g->d3_int_chain, int_chain);
    rc |= upload(&g->d3_orig_of_new, orig_of_new);
    rc |= upload(&g->d3_new_of_orig, new_of_orig);
    rc |= upload(&g->d3_weight, weight3);
    if (rc) return 1;
    CS_CUDA(cudaMalloc(&g->d3_eligible, n));
    g->h3_new_of_orig = new_of_orig;
    g->v3_J = J;
    g->v3_I = I;
    g->v3_ncsec = cnum.size();
    g->v3_ok = true;
    return 0;
}
