#pragma once
#include <stddef.h>
struct gb_ctx;
int comm_allreduce_sum_f64(gb_ctx* ctx, double* buf, size_t count);
int comm_broadcast_f64(gb_ctx* ctx, double* buf, size_t count, int root);
void comm_destroy(gb_ctx* ctx);
