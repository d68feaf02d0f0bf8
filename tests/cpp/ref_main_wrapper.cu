// Drop-in check: the BODY of one of the reference's main() functions (src/bmSparse_SPGEMM.cu:1226-1288 or
// src/bmSparse_SPMV.cu:232-312), unchanged, compiled against this repo's include/ and linked with libbmsparse_b200.so.
// The text of main() is cut out of the reference tree at build time (__graft_entry__.build_reference_mains) into a temporary
// file named by REF_MAIN_BODY -- nothing of it is stored in the repository.  Everything above main() in those files (kernels,
// operators, helper macros) is what this library replaces; only what main() itself needs is provided here.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <string>
#include "bmSpMatrix.h"
#include "reader.h"
#define OUTPUT_TYPE float        // src/bmSparse_SPGEMM.cu:53, src/bmSparse_SPMV.cu:45
using namespace std;
#include REF_MAIN_BODY
