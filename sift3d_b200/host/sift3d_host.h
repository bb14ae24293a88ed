/* sift3d_host.h -- internal declarations of the plain-C host library. */
#ifndef SIFT3D_HOST_H
#define SIFT3D_HOST_H

#include "sift3d_abi.h"
#include "sift3d_cuda.h"

/* host-side Gaussian tap design (init_Gauss_filter, imutil.c:3657-3734);
 * the caller frees g->f.kernel */
int s3dh_gauss_filter(Gauss_filter *g, double sigma, int dim);
int s3dh_gauss_incremental(Gauss_filter *g, double s_cur, double s_next, int dim);

#endif
