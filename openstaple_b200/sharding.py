"""Host-side D3 slab ("salamino") sharding helpers: rank ring, slab ownership and the host-memory
halo exchange, over torch.distributed (gloo or nccl) in place of the reference's MPI.

  ring_neighbours        <- Mpi/multidev.c:60-61  (myrank_L/R = (rank -/+ 1) mod nranks)
  owned_d3_range         <- Mpi/geometry_multidev.h:300-328 (gl_loc_origin_from_rank, D3 only)
  communicate_fermion_borders_hostonly <- Mpi/communications.c:107-156 (host arrays, no device involved)
  communicate_su3_borders_hostonly, communicate_gl3_borders, communicate_tamat_soa_borders, communicate_thmat_soa_borders
                         <- Mpi/communications.c:306-355, 768-784 (rows r0,r1 / all nine entries / packed 3 complex + 2 real)

The device-resident exchange used inside acc_Deo/acc_Doe lives in the CUDA library
(csrc/staple_core.cu: exchange_slices, NCCL over NVLink); these helpers are what a host program uses to
lay out and check its per-rank boxes before uploading them.
"""
from .api import geometry_plan


def ring_neighbours(rank, nranks):
    return (rank + nranks - 1) % nranks, (rank + 1) % nranks


def owned_d3_range(rank, loc_n3):
    """global d3 range [lo, hi) owned by `rank`."""
    return rank * loc_n3, (rank + 1) * loc_n3


def communicate_fermion_borders_hostonly(dist, lnh_fermion, loc_n, halo_width=2, thickness=1):
    """Exchange the first/last interior d3 slice of a HOST vec3_soa (torch CPU complex tensor [3, sizeh])
    with the ring neighbours, in place.  Same offsets, slab sizes and pairing as the reference: per colour,
    slab at send_L goes to rank L (which receives it at recv_R), slab at send_R goes to rank R (recv_L)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return
    p = geometry_plan(loc_n, world, halo_width)
    assert lnh_fermion.shape[-1] == p["sizeh"] and thickness == 1
    L, R = ring_neighbours(rank, world)
    n = p["slab"]
    reqs, recvs = [], []
    for c in range(lnh_fermion.shape[0]):
        row = lnh_fermion[c]
        to_L = row[p["send_L"]:p["send_L"] + n].clone()
        to_R = row[p["send_R"]:p["send_R"] + n].clone()
        from_R, from_L = to_L.new_empty(n), to_L.new_empty(n)
        # tags as in communications.c:76-96: colour c towards L, 3+c towards R
        reqs += [dist.isend(to_L, L, tag=c), dist.isend(to_R, R, tag=3 + c),
                 dist.irecv(from_R, R, tag=c), dist.irecv(from_L, L, tag=3 + c)]
        recvs.append((c, from_R, from_L))
    for r in reqs:
        r.wait()
    for c, from_R, from_L in recvs:
        lnh_fermion[c, p["recv_R"]:p["recv_R"] + n] = from_R
        lnh_fermion[c, p["recv_L"]:p["recv_L"] + n] = from_L


def _exchange_rows(dist, rows, loc_n, halo_width, thickness):
    """Slab exchange of a list of 1-D host tensors (component arrays of length sizeh, complex or real), in place, with the
    reference's offsets (communications.c:51-96): the first `thickness` interior d3 slices go to rank L, which stores them at the
    start of its upper halo; the last `thickness` interior slices go to rank R, which stores them in the slices of its lower
    halo next to the interior.  One message per direction (the reference sends one per component array, same bytes)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return
    import torch
    p = geometry_plan(loc_n, world, halo_width)
    v = p["vol3h"]; off = v * halo_width; n = v * thickness; sizeh = p["sizeh"]
    assert 1 <= thickness <= halo_width and all(r.shape == (sizeh,) for r in rows)
    L, R = ring_neighbours(rank, world)
    to_L = torch.stack([r[off:off + n] for r in rows]).contiguous()
    to_R = torch.stack([r[sizeh - off - n:sizeh - off] for r in rows]).contiguous()
    from_R, from_L = torch.empty_like(to_L), torch.empty_like(to_R)
    flat = lambda t: torch.view_as_real(t).reshape(-1) if t.is_complex() else t.reshape(-1)
    reqs = [dist.isend(flat(to_L), L, tag=0), dist.isend(flat(to_R), R, tag=1),
            dist.irecv(flat(from_R), R, tag=0), dist.irecv(flat(from_L), L, tag=1)]
    for q in reqs:
        q.wait()
    for i, r in enumerate(rows):
        r[sizeh - off:sizeh - off + n] = from_R[i]
        r[off - n:off] = from_L[i]


def communicate_su3_borders_hostonly(dist, lnh_conf, loc_n, thickness, halo_width=2):
    """su3_soa[8] as a host tensor [8, 3, 3, sizeh]: rows r0, r1 of every link (communications.c:319-332)"""
    _exchange_rows(dist, [lnh_conf[k, r, c] for k in range(8) for r in range(2) for c in range(3)], loc_n, halo_width, thickness)


def communicate_gl3_borders(dist, lnh_conf, loc_n, thickness, halo_width=2):
    """gl(3) field in the su3_soa[8] layout: all three rows (communications.c:334-355)"""
    _exchange_rows(dist, [lnh_conf[k, r, c] for k in range(8) for r in range(3) for c in range(3)], loc_n, halo_width, thickness)


def _packed5_rows(field):
    """tamat_soa[8] / thmat_soa[8] as a real host tensor [8, 8, sizeh]: per link three complex arrays (re/im interleaved,
    2*sizeh reals each) followed by two real arrays (struct_c_def.h:45-58)"""
    import torch
    rows = []
    for k in range(8):
        for j in range(3):
            rows.append(torch.view_as_complex(field[k, 2 * j:2 * j + 2].reshape(-1, 2)))
        rows += [field[k, 6], field[k, 7]]
    return rows


def communicate_tamat_soa_borders(dist, lnh_ipdot, loc_n, thickness, halo_width=2):
    """communications.c:776-784"""
    rows = _packed5_rows(lnh_ipdot)
    _exchange_rows(dist, [r for r in rows if r.is_complex()], loc_n, halo_width, thickness)
    _exchange_rows(dist, [r for r in rows if not r.is_complex()], loc_n, halo_width, thickness)


communicate_thmat_soa_borders = communicate_tamat_soa_borders      # same packing, rc00/rc11 in place of ic00/ic11 (:768-774)
