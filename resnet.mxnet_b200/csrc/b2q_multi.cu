// placeholder translation unit (multi-tensor launch: see DESIGN.md "next")
#include "b2q_common.cuh"
