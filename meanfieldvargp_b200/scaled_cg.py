"""
Drop-in for the reference's ``SCG`` class surface (src/numerical/scaled_cg.py:7-409):
``SCG(func, options)(x0) -> (x, fx)``, ``.minimize``, ``.statistics``, ``.max_it``, ``.x_tol``,
``.f_tol``, ``.display``.

The optimisation loop itself runs on the device (mfvgp_scg: iterates, gradients and the
line-search dot products never leave it), which works for ONE kind of objective: the bound
``E_cost`` method of a ``meanfieldvargp_b200.FreeEnergy`` -- exactly what the reference's
``find_minimum`` passes (free_energy.py:820-824).  There is no CPU path in this package: any
other callable is refused with a ``TypeError`` instead of being run in Python.
"""

class SCG(object):

    __slots__ = ("f", "_max_it", "_x_tol", "_f_tol", "display", "_stats")

    def __init__(self, func, *args):
        p_list = args[0] if args else {}                           # scaled_cg.py:36-57
        if not callable(func):
            raise TypeError(f"{self.__class__.__name__}: "
                            f"Function {func} must be callable.")
        self.f = func
        self._max_it = p_list["max_it"] if "max_it" in p_list else 500
        self._x_tol = p_list["x_tol"] if "x_tol" in p_list else 1.0e-6
        self._f_tol = p_list["f_tol"] if "f_tol" in p_list else 1.0e-8
        self.display = p_list["display"] if "display" in p_list else False
        self._stats = None

    @property
    def max_it(self):
        return self._max_it

    @property
    def x_tol(self):
        return self._x_tol

    @property
    def f_tol(self):
        return self._f_tol

    @property
    def statistics(self):                                          # scaled_cg.py:93-105
        if self._stats is None:
            raise NotImplementedError(f" {self.__class__.__name__}:"
                                      f" Stats dictionary has not been created.")
        return self._stats

    def _free_energy(self):
        from .free_energy import FreeEnergy
        owner = getattr(self.f, "__self__", None)
        if isinstance(owner, FreeEnergy) and getattr(self.f, "__name__", "") == "E_cost":
            return owner
        raise TypeError(f"{self.__class__.__name__}: the device loop minimises "
                        f"FreeEnergy.E_cost only (got {self.f}); this package has no CPU path")

    def minimize(self, x0, *args):
        """scaled_cg.py:107-386 -> (x, fx); the statistics dictionary is kept in the object."""
        self._stats = None
        if args:
            raise TypeError(f"{self.__class__.__name__}: E_cost takes no extra arguments")
        fe = self._free_energy()
        x, fx, stats = fe._scg(x0, int(self._max_it), float(self._x_tol), float(self._f_tol))
        if self.display:                                           # :288-300
            for j in range(0, stats["nit"], 50):
                print(f"It= {j:>5}: F(x)= {stats['fx'][j]:.3E} -:- "
                      f"Sum(|Gradients|)= {stats['dfx'][j]:.3E}")
        if stats["nit"] == self._max_it:                           # :376-385
            print(f"SGC: Maximum number of iterations ({self._max_it}) has been reached.")
        self._stats = stats
        return x, fx

    def __call__(self, *args, **kwargs):                           # :388-393
        return self.minimize(*args, **kwargs)

    def __str__(self):
        return f"SCG Id({id(self)}): "\
               f"Function={self.f}, Max-It={self._max_it}, " \
               f"x-tol={self._x_tol}, f-tol={self._f_tol}"
