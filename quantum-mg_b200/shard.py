"""y-slab re-indexing of eo-ordered fields: what a caller uses to cut a global field into the per-rank slabs.

A slab of rows [y0, y0+Yl) of an (X, Y) lattice is itself an (X, Yl) even-odd lattice when y0 is
even (parity is preserved), and every field is two contiguous chunks of the global array, one per
parity (lattice.h:79)."""
import numpy as np


class Slab:
    def __init__(self, X, Y, nranks, rank):
        assert Y % (2 * nranks) == 0
        self.X, self.Y, self.nranks, self.rank = X, Y, nranks, rank
        self.Yl = Y // nranks
        self.y0 = rank * self.Yl
        self.xh = X // 2

    def _chunk(self, par, dof):
        start = ((self.y0 + par * self.Y) * self.xh) * dof
        return slice(start, start + self.Yl * self.xh * dof)

    def take(self, field, dof):
        """Global eo field (V*dof) -> local eo field (X*Yl*dof)."""
        return np.concatenate([field[self._chunk(0, dof)], field[self._chunk(1, dof)]])

    def put(self, field, local, dof):
        h = self.Yl * self.xh * dof
        field[self._chunk(0, dof)] = local[:h]
        field[self._chunk(1, dof)] = local[h:]

    def halo_row(self, field, dof, y_local):
        """Row y_local (-1 or Yl) of the global field as the kernels expect it: (parity, x/2, dof)."""
        y = (self.y0 + y_local) % self.Y
        rows = []
        for par in (0, 1):
            start = ((y + par * self.Y) * self.xh) * dof
            rows.append(field[start:start + self.xh * dof])
        return np.concatenate(rows)
