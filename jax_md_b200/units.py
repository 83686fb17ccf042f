"""Unit systems: drop-in for the reference `jax_md/units.py:68-165` (conversion factors
into the internal eV / Angstrom / amu or kcal/mol / Angstrom / (g/mol) scales; same
CODATA-2014 constants).  Host-side arithmetic only."""
import math

import numpy as np

f64 = np.float64

constants_CONDATA_2014 = {
    '_c': 299792458.0, '_mu0': 4.0e-7 * math.pi, '_Grav': 6.67408e-11,
    '_hplanck': 6.626070040e-34, '_e': 1.6021766208e-19, '_me': 9.10938356e-31,
    '_mp': 1.672621898e-27, '_Nav': 6.022140857e23, '_k': 1.38064852e-23,
    '_amu': 1.660539040e-27,
}


def metal_unit_system(constants=constants_CONDATA_2014):
  """units.py:68-114."""
  Angstrom, eV, amu, charge = 1, 1, 1, 1
  ang = 1e-10
  second = math.sqrt(eV * constants['_e'] / (constants['_amu'] * ang * ang))
  picosecond = 1e-12 * second
  kB = constants['_k'] / constants['_e']
  pascal = eV * ang * ang * ang / constants['_e']
  bar = 1e5 * pascal
  return {
      'mass': f64(amu), 'distance': f64(Angstrom), 'time': f64(picosecond),
      'energy': f64(eV), 'velocity': f64(Angstrom / picosecond),
      'force': f64(eV / Angstrom), 'torque ': f64(eV), 'temperature': f64(kB),
      'pressure': f64(bar), 'charge ': f64(charge),
      'electric field': f64(charge * Angstrom),
  }


def real_unit_system(constants=constants_CONDATA_2014):
  """units.py:117-165."""
  Angstrom, Kcal_mol, amu, charge = 1, 1, 1, 1
  ang = 1e-10
  kcal = 4184.0
  second = math.sqrt(Kcal_mol * kcal / constants['_Nav'] / (constants['_amu'] * ang * ang))
  femtosecond = 1e-15 * second
  kB = constants['_k'] * constants['_Nav'] / kcal
  pascal = Kcal_mol * ang * ang * ang * constants['_Nav'] / kcal
  atm = 101325.0 * pascal
  return {
      'mass': f64(amu), 'distance': f64(Angstrom), 'time': f64(femtosecond),
      'energy': f64(Kcal_mol), 'velocity': f64(Angstrom / femtosecond),
      'force': f64(Kcal_mol / Angstrom), 'torque ': f64(Kcal_mol),
      'temperature': f64(kB), 'pressure': f64(atm), 'charge ': f64(charge),
      'electric field': f64(charge * Angstrom),
  }
