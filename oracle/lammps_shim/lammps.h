// TEST INFRASTRUCTURE: forwards to the single-header LAMMPS shim (see lammps_shim.h).
#include "lammps_shim.h"
