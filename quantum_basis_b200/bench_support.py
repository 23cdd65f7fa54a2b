"""Small helpers shared by bench.py and the sharded bench of dist.py."""


def square_bonds(Lx, Ly):
    """bond list of the Lx x Ly square lattice with periodic boundaries, as examples/trans_absent/latt_square/square_Fermi_Hubbard.cc
    generates it: site = x + y*Lx, one +x and one +y bond per site (duplicates kept)"""
    site = lambda x, y: (x % Lx) + (y % Ly) * Lx   # noqa: E731
    b = []
    for x in range(Lx):
        for y in range(Ly):
            b += [(site(x, y), site(x + 1, y)), (site(x, y), site(x, y + 1))]
    return b
