// see cusp/coo_matrix.h in this directory
#pragma once
#include <cusp/coo_matrix.h>
