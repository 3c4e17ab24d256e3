// TEST INFRASTRUCTURE -- forwards to kokkos_stub.h
#include "kokkos_stub.h"
