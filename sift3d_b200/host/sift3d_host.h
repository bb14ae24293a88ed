/* sift3d_host.h -- internal declarations of the plain-C host library. */
#ifndef SIFT3D_HOST_H
#define SIFT3D_HOST_H

#include "sift3d_abi.h"
#include "sift3d_cuda.h"

/* host-side Gaussian tap design (init_Gauss_filter, imutil.c:3657-3734);
 * the caller frees g->f.kernel */
int s3dh_gauss_filter(Gauss_filter *g, double sigma, int dim);
int s3dh_gauss_incremental(Gauss_filter *g, double s_cur, double s_next, int dim);


/* csv_io.c: write_Mat_rm (imutil.c:1343-1421) with a parallel, printf-exact "%f" formatter;
 * .csv or .csv.gz.  sift3d_b200_format_f writes the characters of printf("%f", v) (at most 352,
 * no terminator) and returns their count. */
int sift3d_b200_write_Mat_rm(const char *path, const Mat_rm *const mat);
int sift3d_b200_format_f(char *dst, double v);

#endif
