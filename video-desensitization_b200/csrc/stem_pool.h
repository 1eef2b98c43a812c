// stem_pool.h — internal interface of the fused conv1 + max-pool kernel (stem_pool.cu), reached through
// vd_conv_plan_create / vd_conv_run with VD_CONV_POOL set.
#pragma once
#include "common.cuh"

struct vd_stem_pool_plan;
int vd_stem_pool_create(const vd_conv_desc_t* d, vd_stem_pool_plan** out);
int vd_stem_pool_run(const vd_stem_pool_plan* pl, cudaStream_t st);
void vd_stem_pool_destroy(vd_stem_pool_plan* pl);
void vd_stem_pool_info(const vd_stem_pool_plan* pl, int* units, int* grid);
void* vd_tensor_map_encoder();      // cuTensorMapEncodeTiled through the runtime's driver entry point, or null

// conv3x3_rows.cu: 3x3 stride-1 pad-1 convolution 64 -> 64 channels with two output rows per accumulator, reached through
// vd_conv_plan_create / vd_conv_run when the descriptor qualifies (VD_CONV_ROWS=0 keeps the generic kernel)
struct vd_conv_rows_plan;
bool vd_conv_rows_eligible(const vd_conv_desc_t* d);
int vd_conv_rows_create(const vd_conv_desc_t* d, vd_conv_rows_plan** out);
int vd_conv_rows_run(const vd_conv_rows_plan* pl, cudaStream_t st);
void vd_conv_rows_destroy(vd_conv_rows_plan* pl);
void vd_conv_rows_info(const vd_conv_rows_plan* pl, int* groups, int* grid);
