// jaxdem_b200 — one (dtype, dim) slice of the cell-list force launchers (see pair.cu: the slices
// exist only to compile in parallel).
#define JDB_PAIR_SLICE_F float
#define JDB_PAIR_SLICE_D 2
#include "pair.cu"
