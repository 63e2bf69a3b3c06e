// jaxdem_b200 — one (dtype, dim) slice of the cell-list force launchers (see pair.cu: the slices
// exist only to compile in parallel).
#define JDB_PAIR_SLICE_F double
#define JDB_PAIR_SLICE_D 3
#include "pair.cu"
