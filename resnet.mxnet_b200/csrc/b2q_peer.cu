// placeholder translation unit (peer-memory threshold exchange: see DESIGN.md "next")
#include "b2q_common.cuh"
