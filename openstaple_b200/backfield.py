"""Host-side producer of the operator's phase field: staggered phases, antiperiodic boundary in T,
imaginary chemical potential and U(1) background E/B field quanta, as ONE angle per link.

    calc_u1_phases  <- OpenAcc/backfield.c:20-187 (calc_u1_phases_unb_no2pi, rebound_u1_phases, mult_u1_phases)
                       and its generated FP32 twin sp_backfield.c (float arithmetic throughout)

The library consumes this array as an input (`backfield` argument of acc_Deo/acc_Doe, `ferm_param.phases`); a host program
that is not the reference's `main` produces it here.  Same operation order as the reference, so the arrays are bit-identical
to its own (tests/test_io_formats.py, against the committed outputs of the reference build and the per-rank boxes of its
two-rank run).  Identity direction map (xmap..tmap = 0,1,2,3) -- the only one for which direction 3, the decomposed one,
is time; k = 2*dir + parity with the parity of the GLOBAL site (backfield.c:88)."""
import numpy as np

from .api import geometry_plan


def calc_u1_phases(loc_n, bf_pars=(0, 0, 0, 0, 0, 0), im_chem_pot=0.0, ferm_charge=0.0, nranks_d3=1, rank=0, halo_width=2,
                   single=False):
    """-> double_soa[8] (float_soa[8] with single=True) as an array [8, sizeh] of angles in radians for this rank's
    local+halo box.  bf_pars = (ex, ey, ez, bx, by, bz) in flux quanta, im_chem_pot = MuOverPiT, ferm_charge = Charge."""
    R = np.float32 if single else np.float64
    p = geometry_plan(loc_n, nranks_d3, halo_width)
    nd0, nd1, nd2, nd3 = p["nd"]; n = p["sizeh"]
    tnx, tny, tnz, tnt = loc_n[0], loc_n[1], loc_n[2], loc_n[3] * nranks_d3
    ex, ey, ez, bx, by, bz = (R(v) for v in bf_pars)
    q = R(ferm_charge); half, one = R(0.5), R(1.0)
    chpotphase = R(im_chem_pot) / R(tnt)                                       # :33
    d3, d2, d1, d0 = np.meshgrid(np.arange(nd3), np.arange(nd2), np.arange(nd1), np.arange(nd0), indexing="ij")
    d0, d1, d2, d3 = (a.reshape(-1) for a in (d0, d1, d2, d3))
    idxh = (d0 + nd0 * (d1 + nd1 * (d2 + nd2 * d3))) // 2                      # snum_acc, geometry_multidev.h:219
    x, y, z, t = d0, d1, d2, d3.copy()
    if nranks_d3 > 1:                                                          # :74-88 (only direction 3 is decomposed)
        t = t + rank * loc_n[3] - p["d3_halo"]
        t = np.where(t > tnt - 1, t - tnt, t); t = np.where(t < 0, t + tnt, t)
    parity = (x + y + z + t) % 2
    f = lambda a: a.astype(R)                                                  # int -> real, as C's usual conversions do
    ph = np.zeros((8, n), R)
    # X-oriented links (:104-118)
    arg = f(z - tnz // 2 + 1) * by / R(tnz * tnx)
    edge = (x + 1 == tnx)
    arg = np.where(edge, (arg - f((y - tny // 2 + 1) * tnx) * bz / R(tnx * tny)) - f((t - tnt // 2 + 1) * tnx) * ex / R(tnx * tnt), arg)
    arg = arg * q
    ph[0 + parity, idxh] = arg
    # Y (:121-133): staggered phase eta_y = (-1)^x
    arg = f(x - tnx // 2 + 1) * bz / R(tnx * tny)
    edge = (y + 1 == tny)
    arg = np.where(edge, (arg - f((z - tnz // 2 + 1) * tny) * bx / R(tny * tnz)) - f((t - tnt // 2 + 1) * tny) * ey / R(tny * tnt), arg)
    arg = arg * q
    arg = np.where(x & 1, arg + half, arg)
    ph[2 + parity, idxh] = arg
    # Z (:136-148): eta_z = (-1)^(x+y)
    arg = f(y - tny // 2 + 1) * bx / R(tny * tnz)
    edge = (z + 1 == tnz)
    arg = np.where(edge, (arg - f((t - tnt // 2 + 1) * tnz) * ez / R(tnz * tnt)) - f((x - tnx // 2 + 1) * tnz) * by / R(tnz * tnx), arg)
    arg = arg * q
    arg = np.where((x + y) & 1, arg + half, arg)
    ph[4 + parity, idxh] = arg
    # T (:151-163): eta_t = (-1)^(x+y+z), chemical potential, antiperiodic boundary on the last time slice
    arg = f(z - tnz // 2 + 1) * ez / R(tnz * tnt)
    arg = arg + f(y - tny // 2 + 1) * ey / R(tny * tnt)
    arg = arg + f(x - tnx // 2 + 1) * ex / R(tnx * tnt)
    arg = arg * q
    arg = np.where((x + y + z) & 1, arg + half, arg)
    arg = arg + chpotphase * half
    arg = np.where(t + 1 == tnt, arg + half, arg)
    ph[6 + parity, idxh] = arg
    # rebound_u1_phases (:166-174) and mult_u1_phases by 2 pi (:176-181)
    while True:
        hi = ph > half
        if not hi.any():
            break
        ph[hi] -= one
    while True:
        lo = ph < -half
        if not lo.any():
            break
        ph[lo] += one
    twopi = R(2 * 3.14159265358979323846) if not single else np.float32(2) * np.float32(3.14159265358979323846)
    return ph * twopi
