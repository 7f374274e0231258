"""Host-side D3 slab ("salamino") sharding helpers: rank ring, slab ownership and the host-memory
halo exchange, over torch.distributed (gloo or nccl) in place of the reference's MPI.

  ring_neighbours        <- Mpi/multidev.c:60-61  (myrank_L/R = (rank -/+ 1) mod nranks)
  owned_d3_range         <- Mpi/geometry_multidev.h:300-328 (gl_loc_origin_from_rank, D3 only)
  communicate_fermion_borders_hostonly <- Mpi/communications.c:107-156 (host arrays, no device involved)

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
