"""On-disk formats and the global <-> rank-local layout either side of the hot path.
Host-side code (numpy), mirroring the reference's own host-side IO so that production configurations can be
loaded straight into the su3_soa[8] arrays the CUDA library reads:

  read_su3_soa_ASCII / print_su3_soa_ASCII            <- OpenAcc/io.c:35-208
  read_su3_soa_ildg_binary / print_su3_soa_ildg_binary <- OpenAcc/io.c:257-551 (LIME records, big-endian doubles,
                                                         sites t,z,y,x slowest->fastest, 4 directions per site)
  print_vec3_soa_ASCII / read_vec3_soa_ASCII          <- DbgTools/dbgtools.c:92-150 (even sites, `re\\tim` lines)
  send_lnh_subconf_to_buffer / recv_loc_subconf_from_buffer, ..._subfermion_... <- Mpi/communications.c:1104-1257
      (global field -> one rank's local+halo box INCLUDING halos, and the interior back)

Layouts: global conf complex128[8, 3, 3, GL_SIZEH] (k = 2*dir + parity), vector complex128[3, GL_SIZEH];
idxh = (x + nx*(y + ny*(z + nz*t)))/2 with the identity direction map (xmap..tmap = 0 1 2 3), which is what the
library's run-time geometry assumes.  Conf files hold all three rows; on reading, the third row is rebuilt as
conj(r0 x r1) exactly as the reference does (`rebuild3row`).
"""
import re
import struct

import numpy as np

ILDG_MAGIC = 0x456789AB
_HDR = struct.Struct(">IHHQ128s")        # ILDG_header (OpenAcc/binary.c:29-36): magic, version, mbme_flag, data_length, type


def _sizeh(dims):
    nx, ny, nz, nt = dims
    return nx * ny * nz * nt // 2


def _cross_conj(a, b, c, d):
    """conj(a*b - c*d) in plain real arithmetic, one rounding per operation like the reference's C99 complex code
    (numpy's own complex multiply may fuse multiply-adds, which changes the last bit and with it the ASCII files)."""
    re = (a.real * b.real - a.imag * b.imag) - (c.real * d.real - c.imag * d.imag)
    im = (a.real * b.imag + a.imag * b.real) - (c.real * d.imag + c.imag * d.real)
    return re - 1j * im


def rebuild3row(conf):
    """third row = conj(r0 x r1), in place (single_types.h:45-51 rebuild3row)."""
    r0, r1 = conf[:, 0], conf[:, 1]
    conf[:, 2, 0] = _cross_conj(r0[:, 1], r1[:, 2], r0[:, 2], r1[:, 1])
    conf[:, 2, 1] = _cross_conj(r0[:, 2], r1[:, 0], r0[:, 0], r1[:, 2])
    conf[:, 2, 2] = _cross_conj(r0[:, 0], r1[:, 1], r0[:, 1], r1[:, 0])
    return conf


# ----------------------------------------------------------------------------- ASCII configurations
def print_su3_soa_ASCII(conf, path, dims, conf_id_iter, beta=0.0, mass=0.0, nflav=0):
    """io.c:35-108: header `nx ny nz nt beta mass NDiffFlavs conf_id`, then for q < 8, i < GL_SIZEH a 3x3 block of
    `(re, im)  ` entries (%.18lf) with the third row rebuilt, one blank line after each matrix."""
    S = _sizeh(dims)
    c = rebuild3row(np.array(conf, dtype=np.complex128, copy=True))
    m = c.transpose(0, 3, 1, 2).reshape(8 * S, 3, 3)            # (q, i) major, then row, col
    with open(path, "w") as f:
        f.write("%d %d %d %d %f %f %d %d\n" % (dims[0], dims[1], dims[2], dims[3], beta, mass, nflav, conf_id_iter))
        out = []
        for mat in m:
            for r in range(3):
                out.append("".join("(%.18f, %.18f)  " % (mat[r, cc].real, mat[r, cc].imag) for cc in range(3)) + "\n")
            out.append("\n")
        f.write("".join(out))


_PAIR = re.compile(r"\(\s*([^,\s]+)\s*,\s*([^)\s]+)\s*\)")


def read_su3_soa_ASCII(path, dims):
    """io.c:111-208 -> (conf[8,3,3,GL_SIZEH], conf_id_iter, header).  Matrices with det = -1 within 0.005 (a
    configuration saved multiplied by the staggered phases) are flipped in sign, as the reference does."""
    S = _sizeh(dims)
    with open(path) as f:
        head = f.readline().split()
        body = f.read()
    nxt, nyt, nzt, ntt = (int(x) for x in head[:4])
    if (nxt, nyt, nzt, ntt) != tuple(dims):
        raise ValueError("configuration dimensions %r not compatible with %r" % ((nxt, nyt, nzt, ntt), tuple(dims)))
    header = dict(beta=float(head[4]), mass=float(head[5]), nflav=int(head[6]))
    vals = np.array(_PAIR.findall(body), dtype=np.float64)
    if vals.shape[0] != 8 * S * 9:
        raise ValueError("not read expected number of entries: %d vs %d" % (vals.shape[0], 8 * S * 9))
    m = (vals[:, 0] + 1j * vals[:, 1]).reshape(8, S, 3, 3)
    det = np.linalg.det(m).real
    m[np.abs(1 + det) < 0.005] *= -1
    return np.ascontiguousarray(m.transpose(0, 2, 3, 1)), int(head[7]), header


# ----------------------------------------------------------------------------- ILDG / LIME configurations
def _site_order(dims):
    """(idxh, parity) of every site in file order t,z,y,x (io.c:553-614, identity direction map)."""
    nx, ny, nz, nt = dims
    t, z, y, x = np.meshgrid(np.arange(nt), np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    idxh = ((x + nx * (y + ny * (z + nz * t))) // 2).ravel()
    parity = ((x + y + z + t) % 2).ravel()
    return idxh, parity


def _pad8(n):
    return 0 if n % 8 == 0 else 8 - n % 8


def print_su3_soa_ildg_binary(conf, path, dims, conf_id_iter, input_file_str=""):
    """io.c:415-551.  Records: ildg-format (xml), MD_traj, input-file, ildg-binary-data, ildg-data-lfn; like the
    reference, data_length counts the padding to 8 bytes."""
    nx, ny, nz, nt = dims
    c = rebuild3row(np.array(conf, dtype=np.complex128, copy=True))
    xml = ("<?xml version=\"1.0\" encoding=\"UTF-8\"?>\n"
           "            <ildgFormat xmlns=\"http://www.lqcd.org/ildg\"\n"
           "            xmlns:xsi=\"http://www.w3.org/2001/XMLSchema-instance\"\n"
           "            xsi:schemaLocation=\"http://www.lqcd.org/ildg filefmt.xsd\">\n"
           "            <version>1.0</version>\n"
           "            <field>su3gauge</field>\n"
           "            <precision>64</precision>\n"
           "            <lx>%d</lx>\n"
           "            <ly>%d</ly>\n"
           "            <lz>%d</lz>\n"
           "            <lt>%d</lt>\n"
           "            </ildgFormat>" % (nx, ny, nz, nt)).encode()

    def record(f, rtype, payload):
        pad = _pad8(len(payload))
        f.write(_HDR.pack(ILDG_MAGIC, 1, 0, len(payload) + pad, rtype.encode()))
        f.write(payload); f.write(b"\0" * pad)

    idxh, parity = _site_order(dims)
    data = np.empty((idxh.size, 4, 3, 3), dtype=np.complex128)
    for d in range(4):
        data[:, d] = c[2 * d + parity, :, :, idxh]           # advanced indexing pairs (k, idxh) per site
    with open(path, "wb") as f:
        record(f, "ildg-format", xml)
        record(f, "MD_traj", ("%d" % conf_id_iter).encode())
        record(f, "input-file", input_file_str.encode())
        record(f, "ildg-binary-data", data.view(np.float64).astype(">f8").tobytes())
        record(f, "ildg-data-lfn", b"")


def read_su3_soa_ildg_binary(path, dims):
    """io.c:257-413 -> (conf[8,3,3,GL_SIZEH], conf_id_iter).  Rows 0,1 are the file's; row 2 is rebuilt from them
    (the reference rebuilds it too but then stores rows 0,1 only, leaving r2 of the destination untouched)."""
    nx, ny, nz, nt = dims
    records = {}
    with open(path, "rb") as f:
        while "ildg-format" not in records or "ildg-binary-data" not in records:
            raw = f.read(_HDR.size)
            if len(raw) != _HDR.size:
                raise ValueError("error in reading ILDG file %s: records missing" % path)
            magic, version, mbme, length, rtype = _HDR.unpack(raw)
            rtype = rtype.split(b"\0", 1)[0].decode()
            records[rtype] = (f.tell(), length)
            f.seek(length + _pad8(length), 1)
        pos, length = records["ildg-format"]
        f.seek(pos); xml = f.read(length).decode(errors="replace")
        got = []
        for tag in ("lx", "ly", "lz", "lt"):
            m = re.search(r"<%s>\s*(\d+)\s*</%s>" % (tag, tag), xml)
            if not m:
                raise ValueError("lx,ly,lz or lt not found in \"ildg-format\"")
            got.append(int(m.group(1)))
        if tuple(got) != tuple(dims):
            raise ValueError("configuration dimensions %r not compatible with %r" % (tuple(got), tuple(dims)))
        conf_id = 1
        if "MD_traj" in records:
            pos, length = records["MD_traj"]
            f.seek(pos); conf_id = int(re.match(rb"\s*(-?\d+)", f.read(length)).group(1))
        pos, _ = records["ildg-binary-data"]
        f.seek(pos)
        nsites = nx * ny * nz * nt
        data = np.frombuffer(f.read(nsites * 4 * 18 * 8), dtype=">f8").astype(np.float64).view(np.complex128)
    data = data.reshape(nsites, 4, 3, 3)
    idxh, parity = _site_order(dims)
    conf = np.zeros((8, 3, 3, nsites // 2), dtype=np.complex128)
    for d in range(4):
        conf[2 * d + parity, :, :, idxh] = data[:, d]
    return rebuild3row(conf), conf_id


# ----------------------------------------------------------------------------- ASCII vectors (debug dumps)
def _even_site_order(dims):
    nx, ny, nz, nt = dims
    t, z, y, xh = np.meshgrid(np.arange(nt), np.arange(nz), np.arange(ny), np.arange(nx // 2), indexing="ij")
    x = 2 * xh + ((y + z + t) & 1)
    return ((x + nx * (y + ny * (z + nz * t))) // 2).ravel()


def print_vec3_soa_ASCII(vec, path, dims):
    """dbgtools.c:92-116 save_gl_fermion: per even site in order t,z,y,x/2 three lines `re\\tim` (%.18lf).
    (Format restated from the source; the reference's dbgtools is not part of the oracle build, so this pair is
    checked by round trip only.)"""
    order = _even_site_order(dims)
    v = np.asarray(vec)[:, order].T.reshape(-1)
    with open(path, "w") as f:
        f.write("".join("%.18f\t%.18f\n" % (z.real, z.imag) for z in v))


def read_vec3_soa_ASCII(path, dims):
    order = _even_site_order(dims)
    vals = np.loadtxt(path, dtype=np.float64).reshape(-1, 3, 2)
    vec = np.zeros((3, _sizeh(dims)), dtype=np.complex128)
    vec[:, order] = (vals[..., 0] + 1j * vals[..., 1]).T
    return vec


# ----------------------------------------------------------------------------- global <-> rank-local boxes
def _box(loc_n, nranks, halo_width):
    from .api import geometry_plan
    p = geometry_plan(loc_n, nranks, halo_width)
    return p["nd"][3], p["vol3h"], p["sizeh"], p["d3_halo"]


def _gl_d3(d3, rank, loc_n3, d3_halo, gl_n3):
    return (d3 + rank * loc_n3 - d3_halo) % gl_n3       # geometry_multidev.h:246-262, periodic through the last rank


def send_lnh_subfield_to_buffer(gl, target_rank, loc_n, nranks, halo_width=2):
    """communications.c:1104-1148 (links) / :1151-1190 (fermions): the target rank's local+halo box, HALOS INCLUDED,
    cut out of a global field whose last axis is GL_SIZEH.  Only D3 is decomposed, so every local d3 slice is one
    contiguous block of vol3h elements of the global slice (d3 + rank*LOC_N3 - D3_HALO) mod GL_N3; site parities are
    preserved because LOC_N3 and the halo are even (the library rejects the other cases)."""
    nd3, vol3h, sizeh, d3_halo = _box(loc_n, nranks, halo_width)
    gl = np.asarray(gl)
    gl_n3 = loc_n[3] * nranks
    assert gl.shape[-1] == vol3h * gl_n3
    out = np.empty(gl.shape[:-1] + (sizeh,), dtype=gl.dtype)
    for d3 in range(nd3):
        g3 = _gl_d3(d3, target_rank, loc_n[3], d3_halo, gl_n3)
        out[..., d3 * vol3h:(d3 + 1) * vol3h] = gl[..., g3 * vol3h:(g3 + 1) * vol3h]
    return out


def recv_loc_subfield_from_buffer(gl, lnh, rank, loc_n, nranks, halo_width=2):
    """communications.c:1149-1257: copy the INTERIOR of a rank's box back into the global field (in place)."""
    nd3, vol3h, sizeh, d3_halo = _box(loc_n, nranks, halo_width)
    gl_n3 = loc_n[3] * nranks
    for d3 in range(d3_halo, d3_halo + loc_n[3]):
        g3 = _gl_d3(d3, rank, loc_n[3], d3_halo, gl_n3)
        gl[..., g3 * vol3h:(g3 + 1) * vol3h] = lnh[..., d3 * vol3h:(d3 + 1) * vol3h]
    return gl


send_lnh_subconf_to_buffer = send_lnh_subfield_to_buffer
send_lnh_subfermion_to_buffer = send_lnh_subfield_to_buffer
recv_loc_subconf_from_buffer = recv_loc_subfield_from_buffer
recv_loc_subfermion_from_buffer = recv_loc_subfield_from_buffer
