"""Unit-conversion scalars used on the TDSE path (host scalars only).

Mirrors the call signatures of the reference's `richmol/convert_units.py:5-123`
(`AUpol_x_Vm_to_invcm` :102, `AUdip_x_Vm_to_invcm` :83, `Debye_x_Vm_to_invcm` :41,
`J_to_invcm` :11, `MHz_to_invcm` :5): called without arguments each returns the factor,
called with arguments it returns a tuple of the scaled arguments.
All factors come from the CODATA values shipped in `scipy.constants`.
"""
from scipy import constants as _c

_H = _c.value("Planck constant")
_C = _c.value("speed of light in vacuum")


def _apply(factor, args):
    if not args:
        return factor
    out = []
    for a in args:
        try:
            out.append([x * factor for x in a])
        except TypeError:
            out.append(a * factor)
    return tuple(out)


def MHz_to_invcm(*args):
    return _apply(1 / _C * 1e4, args)


def J_to_invcm(*args):
    return _apply(1 / (_H * 1e2 * _C), args)


def Debye_to_au(*args):
    f = 1e-21 / _C / _c.value("elementary charge") / _c.value("Bohr radius")
    return _apply(f, args)


def Debye_to_si(*args):
    return _apply(1e-21 / _C, args)


def AUdip_x_Vm_to_invcm(*args):
    f = _c.value("atomic unit of electric dipole mom.") / (_H * _C) / 1e2
    return _apply(f, args)


def Debye_x_Vm_to_invcm(*args):
    return _apply(AUdip_x_Vm_to_invcm() * Debye_to_au(), args)


def AUpol_x_Vm_to_invcm(*args):
    f = _c.value("atomic unit of electric polarizability") / (_H * _C) / 1e2
    return _apply(f, args)


def AUdip_x_Vm_to_MHz(*args):
    return _apply(AUdip_x_Vm_to_invcm() / MHz_to_invcm(), args)


def Debye_x_Vm_to_MHz(*args):
    return _apply(Debye_x_Vm_to_invcm() / MHz_to_invcm(), args)
