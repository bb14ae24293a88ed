/* sift3d_api.c -- drop-in libsift3D.so for NVIDIA B200: the reference's public C
 * API (sift3d/sift.h:19-108) on top of the device engine (include/sift3d_cuda.h).
 *
 * What stays on the host, in plain C, mirrors what the reference keeps cheap and
 * scalar: parameter validation (sift.c:514-580), pyramid geometry
 * (resize_Pyramid imutil.c:3858, set_scales_Pyramid imutil.c:3957), Gaussian tap
 * design in f64 (init_Gauss_filter imutil.c:3657), the icosahedron table
 * (init_geometry sift.c:215), the Keypoint/Descriptor stores and their converters.
 * Everything per-voxel or per-keypoint runs in libsift3d_cuda.so.  There is no
 * CPU fallback: if the engine cannot be created the hot calls return
 * SIFT3D_FAILURE with a message on stderr.
 *
 * Each SIFT3D object owns one engine, found through a small handle stored in the
 * struct's (otherwise dead) OpenCL field `kernels.downsample_2`, so the struct
 * keeps its exact reference layout and can be copied by value.
 */
#define _GNU_SOURCE
#include "sift3d_host.h"

#include <float.h>
#include <getopt.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- exported constants of the reference (sift.c:34-58) ---------------------- */
const double peak_thresh_default = 0.1;
const int num_kp_levels_default = 3;
const double corner_thresh_default = 0.4;
const double sigma_n_default = 1.15;
const double sigma0_default = 1.6;
const char opt_peak_thresh[] = "peak_thresh";
const char opt_corner_thresh[] = "corner_thresh";
const char opt_num_kp_levels[] = "num_kp_levels";
const char opt_sigma_n[] = "sigma_n";
const char opt_sigma0[] = "sigma0";
const double ori_sig_fctr = 1.5;   /* sift.c:51 */
const double ori_rad_fctr = 3.0;   /* sift.c:52 */
const double desc_sig_fctr = 7.071067812;
const double desc_rad_fctr = 2.0;  /* sift.c:54 */
const double gr = 1.6180339887;
/* internal parameters the reference also exports as data symbols (sift.c:48-55); the device
 * kernels carry the same values (keypoint.cu) */
const double max_eig_ratio = 0.90;
const double ori_grad_thresh = 1E-10;
const double bary_eps = FLT_EPSILON * 1E1;
const double trunc_thresh = 0.2f * 128.0f / 768;

#define ERR(...) fprintf(stderr, __VA_ARGS__)
#define MAXV(a, b) ((a) > (b) ? (a) : (b))
#define MINV(a, b) ((a) < (b) ? (a) : (b))

/* ============================================================ engine registry */

typedef struct Slot {
    s3d_engine *eng;
    int have_image;        /* an image has been set (reference: sift3d->im.data != NULL) */
    double resize_units[3]; /* units at the last resize_SIFT3D (see level_geometry) */
    int ncand;
} Slot;

static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static pthread_mutex_t g_util_lock = PTHREAD_MUTEX_INITIALIZER; /* serialises the shared engine */
static Slot **g_slots = NULL; /* stable Slot addresses; index 0 = "no engine" */
static int g_nslots = 0;

static int slot_new(void)
{
    int id = 0, i;
    Slot *fresh = (Slot *)calloc(1, sizeof(Slot));
    if (!fresh) return 0;
    pthread_mutex_lock(&g_lock);
    for (i = 1; i < g_nslots && !id; i++)
        if (!g_slots[i]) id = i;
    if (!id) {
        const int n = g_nslots ? 2 * g_nslots : 16;
        Slot **p = (Slot **)realloc(g_slots, n * sizeof(Slot *));
        if (p) {
            for (i = g_nslots; i < n; i++) p[i] = NULL;
            id = g_nslots ? g_nslots : 1;
            g_slots = p;
            g_nslots = n;
        }
    }
    if (id)
        g_slots[id] = fresh;
    else
        free(fresh);
    pthread_mutex_unlock(&g_lock);
    return id;
}

static Slot *slot_get(const SIFT3D *s)
{
    const int id = s->kernels.downsample_2;
    Slot *r = NULL;
    pthread_mutex_lock(&g_lock);
    if (id > 0 && id < g_nslots) r = g_slots[id];
    pthread_mutex_unlock(&g_lock);
    return r;
}

static void slot_release(const SIFT3D *s)
{
    const int id = s->kernels.downsample_2;
    Slot *r = NULL;
    pthread_mutex_lock(&g_lock);
    if (id > 0 && id < g_nslots) {
        r = g_slots[id];
        g_slots[id] = NULL;
    }
    pthread_mutex_unlock(&g_lock);
    if (r) {
        if (r->eng) s3d_engine_destroy(r->eng);
        free(r);
    }
}

/* Lazily create the device engine (so init_SIFT3D works without a GPU, like the
 * reference's option parsing does). */
static s3d_engine *engine_of(const SIFT3D *s)
{
    Slot *sl = slot_get(s);
    if (!sl) {
        ERR("sift3d_b200: SIFT3D struct was not initialised with init_SIFT3D \n");
        return NULL;
    }
    if (!sl->eng) {
        float v[ICOS_NFACES * 9];
        int idx[ICOS_NFACES * 3], i, j;
        if (s3d_engine_create(&sl->eng, -1)) {
            ERR("sift3d_b200: cannot create the CUDA engine: %s \n", s3d_last_create_error());
            sl->eng = NULL;
            return NULL;
        }
        for (i = 0; i < ICOS_NFACES; i++)
            for (j = 0; j < 3; j++) {
                v[9 * i + 3 * j + 0] = s->mesh.tri[i].v[j].x;
                v[9 * i + 3 * j + 1] = s->mesh.tri[i].v[j].y;
                v[9 * i + 3 * j + 2] = s->mesh.tri[i].v[j].z;
                idx[3 * i + j] = s->mesh.tri[i].idx[j];
            }
        if (s3d_set_mesh(sl->eng, v, idx)) {
            s3d_engine_destroy(sl->eng);
            sl->eng = NULL;
            return NULL;
        }
    }
    return sl->eng;
}

void *sift3d_b200_engine(const SIFT3D *sift3d) { return engine_of(sift3d); }

int sift3d_b200_num_candidates(const SIFT3D *sift3d)
{
    Slot *sl = slot_get(sift3d);
    return sl ? sl->ncand : -1;
}

/* ============================================================ small host helpers */

static void *safe_realloc(void *ptr, size_t size)
{ /* SIFT3D_safe_realloc semantics, imutil.c:247-260 */
    void *ret;
    if (size == 0 || (ret = realloc(ptr, size)) == NULL) {
        free(ptr);
        return NULL;
    }
    return ret;
}

static size_t mat_type_size(Mat_rm_type t)
{
    return t == SIFT3D_DOUBLE ? sizeof(double) : (t == SIFT3D_FLOAT ? sizeof(float) : sizeof(int));
}

/* resize_Mat_rm semantics (imutil.c:844-897) */
static int mat_resize(Mat_rm *m)
{
    const size_t total = mat_type_size(m->type) * (size_t)(m->num_rows * m->num_cols);
    if (m->type != SIFT3D_DOUBLE && m->type != SIFT3D_FLOAT && m->type != SIFT3D_INT) {
        ERR("resize_Mat_rm: unknown type! \n");
        return SIFT3D_FAILURE;
    }
    if (total == m->size) return SIFT3D_SUCCESS;
    m->size = total;
    if (m->static_mem) {
        ERR("resize_Mat_rm: illegal re-allocation of static matrix \n");
        return SIFT3D_FAILURE;
    }
    if (total == 0) {
        free(m->u.data_double);
        m->u.data_double = NULL;
        return SIFT3D_SUCCESS;
    }
    if ((m->u.data_double = (double *)safe_realloc(m->u.data_double, total)) == NULL) {
        m->size = 0;
        return SIFT3D_FAILURE;
    }
    return SIFT3D_SUCCESS;
}

static void mat_init_empty(Mat_rm *m, Mat_rm_type t)
{
    memset(m, 0, sizeof(*m));
    m->type = t;
}

static void image_blank(Image *im)
{ /* init_im, imutil.c:3626-3648 */
    memset(im, 0, sizeof(*im));
    im->ux = im->uy = im->uz = 1.0;
    im->s = -1.0;
}

static void image_default_stride(Image *im)
{ /* im_default_stride, imutil.c:1453-1466 */
    im->xs = (size_t)im->nc;
    im->ys = im->xs * (size_t)im->nx;
    im->zs = im->ys * (size_t)im->ny;
}

/* init_Gauss_filter, imutil.c:3657-3710: taps designed in f64, normalised in f32 */
int s3dh_gauss_filter(Gauss_filter *g, double sigma, int dim)
{
    const int hw = sigma > 0 ? MAXV((int)ceil(sigma * 3.0), 1) : 1;
    const int width = 2 * hw + 1;
    float acc = 0, *k;
    int i;
    if ((k = (float *)malloc(width * sizeof(float))) == NULL) return SIFT3D_FAILURE;
    for (i = 0; i < width; i++) {
        double x = (double)i - hw;
        x /= sigma + DBL_EPSILON;
        k[i] = (float)exp(-0.5 * x * x);
        acc += k[i];
    }
    for (i = 0; i < width; i++) k[i] /= acc;
    g->sigma = sigma;
    g->f.cl_apply_unrolled = 0;
    g->f.kernel = k;
    g->f.dim = dim;
    g->f.width = width;
    g->f.symmetric = SIFT3D_TRUE;
    return SIFT3D_SUCCESS;
}

/* init_Gauss_incremental_filter, imutil.c:3713-3734 */
int s3dh_gauss_incremental(Gauss_filter *g, double s_cur, double s_next, int dim)
{
    if (s_cur > s_next) {
        ERR("init_Gauss_incremental_filter: s_cur (%f) > s_next (%f) \n", s_cur, s_next);
        return SIFT3D_FAILURE;
    }
    return s3dh_gauss_filter(g, sqrt(s_next * s_next - s_cur * s_cur), dim);
}

static void gss_free(GSS_filters *gss)
{
    int i;
    if (gss->num_filters < 1) return;
    free(gss->first_gauss.f.kernel);
    gss->first_gauss.f.kernel = NULL;
    for (i = 0; i < gss->num_filters; i++) free(gss->gauss_octave[i].f.kernel);
    free(gss->gauss_octave);
    gss->gauss_octave = NULL;
    gss->num_filters = -1;
}

#define PYR_LEVEL(pyr, o, s) \
    ((pyr)->levels + ((o) - (pyr)->first_octave) * (pyr)->num_levels + ((s) - (pyr)->first_level))

/* set_scales_Pyramid, imutil.c:3957-3992 */
static int pyr_set_scales(double sigma0, double sigma_n, Pyramid *pyr)
{
    int o, s;
    for (o = pyr->first_octave; o < pyr->first_octave + pyr->num_octaves; o++)
        for (s = pyr->first_level; s < pyr->first_level + pyr->num_levels; s++) {
            const double scale = sigma0 * pow(2.0, o + (double)s / pyr->num_kp_levels);
            if (o == pyr->first_octave && s == pyr->first_level && scale < sigma_n) {
                ERR("set_scales_Pyramid: sigma_n too large for these settings. "
                    "Max allowed: %f \n", scale - DBL_EPSILON);
                return SIFT3D_FAILURE;
            }
            PYR_LEVEL(pyr, o, s)->s = scale;
        }
    pyr->sigma0 = sigma0;
    pyr->sigma_n = sigma_n;
    return SIFT3D_SUCCESS;
}

/* host copies of level data exist only after sift3d_b200_materialize_pyramids */
static void pyr_free_host_data(Pyramid *pyr, int total)
{
    int i;
    if (!pyr->levels) return;
    for (i = 0; i < total; i++) {
        free(pyr->levels[i].data);
        pyr->levels[i].data = NULL;
        pyr->levels[i].size = 0;
    }
}

/* resize_Pyramid, imutil.c:3858-3947 -- metadata only: the voxel data of every level lives
 * in HBM.  Image.data stays NULL until the caller asks for host copies
 * (sift3d_b200_materialize_pyramids, or $SIFT3D_HOST_PYRAMID=1; see sift3d_b200_fetch_level). */
static int pyr_resize(const Image *im, int have_image, int first_level, unsigned num_kp_levels,
                      unsigned num_levels, int first_octave, unsigned num_octaves, Pyramid *pyr)
{
    const int total = (int)(num_levels * num_octaves);
    const int pyr_old_total = pyr->levels ? pyr->num_levels * pyr->num_octaves : 0;
    double units[3];
    int dims[3], i, o, s;
    if (num_levels < num_kp_levels) {
        ERR("resize_Pyramid: num_levels (%u) < num_kp_levels (%d)", num_levels, num_kp_levels);
        return SIFT3D_FAILURE;
    }
    pyr->first_level = first_level;
    pyr->num_kp_levels = (int)num_kp_levels;
    pyr->first_octave = first_octave;
    pyr->num_octaves = (int)num_octaves;
    pyr->num_levels = (int)num_levels;
    pyr_free_host_data(pyr, pyr_old_total);
    if (total == 0) {
        free(pyr->levels);
        pyr->levels = NULL;
        return SIFT3D_SUCCESS;
    }
    if ((pyr->levels = (Image *)safe_realloc(pyr->levels, total * sizeof(Image))) == NULL)
        return SIFT3D_FAILURE;
    for (i = 0; i < total; i++) image_blank(&pyr->levels[i]);
    if (!have_image) return SIFT3D_SUCCESS;
    for (i = 0; i < 3; i++) {
        dims[i] = (int)((double)(&im->nx)[i] * pow(2.0, -first_octave));
        units[i] = (&im->ux)[i] * pow(2.0, -first_octave);
    }
    for (o = first_octave; o < first_octave + (int)num_octaves; o++) {
        for (s = first_level; s < first_level + (int)num_levels; s++) {
            Image *l = PYR_LEVEL(pyr, o, s);
            l->nx = dims[0];
            l->ny = dims[1];
            l->nz = dims[2];
            l->ux = units[0];
            l->uy = units[1];
            l->uz = units[2];
            l->nc = im->nc;
            image_default_stride(l);
        }
        for (i = 0; i < 3; i++) {
            dims[i] /= 2;
            units[i] *= 2;
        }
    }
    return pyr_set_scales(pyr->sigma0, pyr->sigma_n, pyr);
}

/* make_gss, imutil.c:3752-3802: filters from the octave-0 scales only */
static int gss_make(GSS_filters *gss, const Pyramid *pyr)
{
    const int num_filters = pyr->num_levels - 1;
    const int first_level = pyr->first_level;
    int s;
    if (num_filters < 1) {
        ERR("make_gss: pyr has only %d levels, must have at least 2", pyr->num_levels);
        return SIFT3D_FAILURE;
    }
    gss_free(gss);
    if ((gss->gauss_octave = (Gauss_filter *)calloc(num_filters, sizeof(Gauss_filter))) == NULL)
        return SIFT3D_FAILURE;
    gss->num_filters = num_filters;
    gss->first_level = first_level;
    if (s3dh_gauss_incremental(&gss->first_gauss, pyr->sigma_n,
                               PYR_LEVEL(pyr, pyr->first_octave, first_level)->s, 3))
        return SIFT3D_FAILURE;
    for (s = first_level; s < first_level + pyr->num_levels - 1; s++)
        if (s3dh_gauss_incremental(&gss->gauss_octave[s - first_level],
                                   PYR_LEVEL(pyr, pyr->first_octave, s)->s,
                                   PYR_LEVEL(pyr, pyr->first_octave, s + 1)->s, 3))
            return SIFT3D_FAILURE;
    return SIFT3D_SUCCESS;
}

/* init_geometry, sift.c:215-326, including its quirk: every face takes the
 * "swap" branch, which exchanges the vertex vectors v[0]<->v[1] but leaves
 * idx[0], idx[1] in place. */
static int mesh_build(Mesh *mesh)
{
    const float vert[ICOS_NVERT][3] = {{0, 1, gr},  {0, -1, gr}, {0, 1, -gr},  {0, -1, -gr},
                                       {1, gr, 0},  {-1, gr, 0}, {1, -gr, 0},  {-1, -gr, 0},
                                       {gr, 0, 1},  {-gr, 0, 1}, {gr, 0, -1},  {-gr, 0, -1}};
    const int faces[ICOS_NFACES][3] = {{0, 1, 8},  {0, 8, 4},  {0, 4, 5},  {0, 5, 9},  {0, 9, 1},
                                       {1, 6, 8},  {8, 6, 10}, {8, 10, 4}, {4, 10, 2}, {4, 2, 5},
                                       {5, 2, 11}, {5, 11, 9}, {9, 11, 7}, {9, 7, 1},  {1, 7, 6},
                                       {3, 6, 7},  {3, 7, 11}, {3, 11, 2}, {3, 2, 10}, {3, 10, 6}};
    int i, j;
    if ((mesh->tri = (Tri *)safe_realloc(NULL, ICOS_NFACES * sizeof(Tri))) == NULL)
        return SIFT3D_FAILURE;
    mesh->num = ICOS_NFACES;
    for (i = 0; i < ICOS_NFACES; i++) {
        Cvec *const v = mesh->tri[i].v;
        Cvec a, b, n;
        for (j = 0; j < 3; j++) {
            const int id = faces[i][j];
            float mag;
            mesh->tri[i].idx[j] = id;
            v[j].x = vert[id][0];
            v[j].y = vert[id][1];
            v[j].z = vert[id][2];
            mag = sqrtf(v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z);
            /* the reference's SIFT3D_CVEC_SCALE(v + j, 1.0f / mag) expands textually to
             * x * 1.0f / mag (sift.c:295, immacros.h:291-295) */
            v[j].x = v[j].x * 1.0f / mag;
            v[j].y = v[j].y * 1.0f / mag;
            v[j].z = v[j].z * 1.0f / mag;
        }
        a.x = v[2].x - v[1].x, a.y = v[2].y - v[1].y, a.z = v[2].z - v[1].z;
        b.x = v[1].x - v[0].x, b.y = v[1].y - v[0].y, b.z = v[1].z - v[0].z;
        n.x = a.y * b.z - a.z * b.y;
        n.y = a.z * b.x - a.x * b.z;
        n.z = a.x * b.y - a.y * b.x;
        if (n.x * v[0].x + n.y * v[0].y + n.z * v[0].z < 0) {
            const Cvec t = v[0];
            v[0] = v[1];
            v[1] = t;
        }
    }
    return SIFT3D_SUCCESS;
}

/* ============================================================ stores */

void init_Keypoint_store(Keypoint_store *const kp)
{
    kp->slab.buf_size = kp->slab.num = 0;
    kp->slab.buf = NULL;
    kp->buf = NULL;
}

int init_Keypoint(Keypoint *const key)
{ /* R aliases r_data, static memory (sift.c:406-410, imutil.c:655-678) */
    key->R.type = SIFT3D_FLOAT;
    key->R.num_rows = key->R.num_cols = IM_NDIMS;
    key->R.size = IM_NDIMS * IM_NDIMS * sizeof(float);
    key->R.u.data_float = key->r_data;
    key->R.static_mem = SIFT3D_TRUE;
    return SIFT3D_SUCCESS;
}

int resize_Keypoint_store(Keypoint_store *const kp, const size_t num)
{ /* sift.c:417-436 with SIFT3D_RESIZE_SLAB (immacros.h:199-222): 500-element steps */
    void *const buf_old = kp->slab.buf;
    const size_t slab_len = 500;
    const size_t slabs_new = (num + slab_len - 1) / slab_len;
    const size_t size_new = slabs_new * slab_len * sizeof(Keypoint);
    size_t i;
    if (size_new != kp->slab.buf_size) {
        if (size_new == 0) {
            free(kp->slab.buf);
            kp->slab.buf = NULL;
        } else if ((kp->slab.buf = safe_realloc(kp->slab.buf, size_new)) == NULL) {
            return SIFT3D_FAILURE;
        }
        kp->slab.buf_size = size_new;
    }
    kp->slab.num = num;
    kp->buf = (Keypoint *)kp->slab.buf;
    if (buf_old != kp->slab.buf)
        for (i = 0; i < kp->slab.num; i++) init_Keypoint(kp->buf + i);
    return SIFT3D_SUCCESS;
}

int copy_Keypoint(const Keypoint *const src, Keypoint *const dst)
{ /* sift.c:439-451 */
    dst->xd = src->xd;
    dst->yd = src->yd;
    dst->zd = src->zd;
    dst->sd = src->sd;
    dst->o = src->o;
    dst->s = src->s;
    dst->R.type = src->R.type;
    dst->R.num_rows = src->R.num_rows;
    dst->R.num_cols = src->R.num_cols;
    if (mat_resize(&dst->R)) return SIFT3D_FAILURE;
    memmove(dst->R.u.data_double, src->R.u.data_double, src->R.size);
    return SIFT3D_SUCCESS;
}

void cleanup_Keypoint_store(Keypoint_store *const kp) { free(kp->slab.buf); }

void init_SIFT3D_Descriptor_store(SIFT3D_Descriptor_store *const desc) { desc->buf = NULL; }

void cleanup_SIFT3D_Descriptor_store(SIFT3D_Descriptor_store *const desc) { free(desc->buf); }

static int desc_store_resize(SIFT3D_Descriptor_store *const desc, const int num)
{ /* sift.c:476-491 */
    if (num < 1) {
        ERR("resize_SIFT3D_Descriptor_store: invalid size: %d", num);
        return SIFT3D_FAILURE;
    }
    if ((desc->buf = (SIFT3D_Descriptor *)safe_realloc(
             desc->buf, (size_t)num * sizeof(SIFT3D_Descriptor))) == NULL)
        return SIFT3D_FAILURE;
    desc->num = (size_t)num;
    return SIFT3D_SUCCESS;
}

/* ============================================================ parameters */

int set_peak_thresh_SIFT3D(SIFT3D *const sift3d, const double peak_thresh)
{ /* sift.c:514-524 */
    if (peak_thresh <= 0.0 || peak_thresh > 1) {
        ERR("SIFT3D peak_thresh must be in the interval (0, 1]. Provided: %f \n", peak_thresh);
        return SIFT3D_FAILURE;
    }
    sift3d->peak_thresh = peak_thresh;
    return SIFT3D_SUCCESS;
}

int set_corner_thresh_SIFT3D(SIFT3D *const sift3d, const double corner_thresh)
{ /* sift.c:527-538 */
    if (corner_thresh < 0.0 || corner_thresh > 1.0) {
        ERR("SIFT3D corner_thresh must be in the interval [0, 1]. Provided: %f \n", corner_thresh);
        return SIFT3D_FAILURE;
    }
    sift3d->corner_thresh = corner_thresh;
    return SIFT3D_SUCCESS;
}

/* Geometry the reference would hold in its pyramids at the next build_gpyr.
 * QUIRK kept on purpose: dims and units of every level are only recomputed by
 * resize_Pyramid, which set_im_SIFT3D calls when the DIMENSIONS change
 * (sift.c:906-910).  A new image with equal dims but different units updates
 * octave 0 (blur outputs copy the source's units, imutil.c:3478) while octaves
 * >= 1 keep the units of the last resize (im_downsample_2x does not touch them,
 * imutil.c:1742-1768). */
static int push_geometry_ex(SIFT3D *s, Slot *sl, const int *zsplit, s3d_comm *comm);
static int push_geometry(SIFT3D *s, Slot *sl) { return push_geometry_ex(s, sl, NULL, NULL); }

/* zsplit/comm != NULL: Z-slab tiling (the geometry pushed is the GLOBAL one; the engine
 * derives this rank's share, see include/sift3d_cuda.h) */
static int push_geometry_ex(SIFT3D *s, Slot *sl, const int *zsplit, s3d_comm *comm)
{
    const Pyramid *g = &s->gpyr, *d = &s->dog;
    const int ng = g->num_levels * g->num_octaves, nd = d->num_levels * d->num_octaves;
    s3d_geom *gg, *dd;
    s3d_filter first, *oct;
    int i, o, rc;
    if (ng < 1) return SIFT3D_SUCCESS;
    gg = (s3d_geom *)calloc(ng + nd, sizeof(s3d_geom));
    oct = (s3d_filter *)calloc(s->gss.num_filters, sizeof(s3d_filter));
    dd = gg + ng;
    for (i = 0; i < ng + nd; i++) {
        const Image *l = i < ng ? &g->levels[i] : &d->levels[i - ng];
        s3d_geom *q = &gg[i];
        o = (i < ng ? i / g->num_levels : (i - ng) / d->num_levels);
        q->nx = l->nx;
        q->ny = l->ny;
        q->nz = l->nz;
        q->scale = l->s;
        if (o == 0) {
            q->ux = s->im.ux, q->uy = s->im.uy, q->uz = s->im.uz;
        } else {
            q->ux = ldexp(sl->resize_units[0], o);
            q->uy = ldexp(sl->resize_units[1], o);
            q->uz = ldexp(sl->resize_units[2], o);
        }
    }
    /* mirror into the host metadata so callers reading the Pyramid see the truth */
    for (i = 0; i < ng; i++)
        g->levels[i].ux = gg[i].ux, g->levels[i].uy = gg[i].uy, g->levels[i].uz = gg[i].uz;
    for (i = 0; i < nd; i++)
        d->levels[i].ux = dd[i].ux, d->levels[i].uy = dd[i].uy, d->levels[i].uz = dd[i].uz;
    first.taps = s->gss.first_gauss.f.kernel;
    first.width = s->gss.first_gauss.f.width;
    for (i = 0; i < s->gss.num_filters; i++) {
        oct[i].taps = s->gss.gauss_octave[i].f.kernel;
        oct[i].width = s->gss.gauss_octave[i].f.width;
    }
    rc = s3d_pyramid_filters(sl->eng, &first, oct, s->gss.num_filters);
    if (!rc)
        rc = comm ? s3d_slab_pyramid_resize(sl->eng, comm, g->num_octaves, g->num_kp_levels, gg,
                                            dd, zsplit)
                  : s3d_pyramid_resize(sl->eng, g->num_octaves, g->num_kp_levels, gg, dd);
    free(gg);
    free(oct);
    return rc ? SIFT3D_FAILURE : SIFT3D_SUCCESS;
}

/* resize_SIFT3D, sift.c:938-986 */
static int resize_all(SIFT3D *const s, const int num_kp_levels)
{
    Slot *sl = slot_get(s);
    const int have = sl && sl->have_image;
    const unsigned num_dog_levels = num_kp_levels + 2, num_gpyr_levels = num_dog_levels + 1;
    int num_octaves = 0;
    if (have) {
        const int last_octave =
            (int)log2((double)MINV(MINV(s->im.nx, s->im.ny), s->im.nz)) - 3;
        if (last_octave < 0) {
            ERR("resize_SIFT3D: input image is too small: must have at least 8 voxels in "
                "each dimension \n");
            return SIFT3D_FAILURE;
        }
        num_octaves = last_octave + 1;
    }
    if (pyr_resize(&s->im, have, -1, num_kp_levels, num_gpyr_levels, 0, num_octaves, &s->gpyr) ||
        pyr_resize(&s->im, have, -1, num_kp_levels, num_dog_levels, 0, num_octaves, &s->dog))
        return SIFT3D_FAILURE;
    if (!have) return SIFT3D_SUCCESS;
    sl->resize_units[0] = s->im.ux;
    sl->resize_units[1] = s->im.uy;
    sl->resize_units[2] = s->im.uz;
    return gss_make(&s->gss, &s->gpyr);
}

/* set_scales_SIFT3D, sift.c:916-934 */
static int set_scales(SIFT3D *const s, const double sigma0, const double sigma_n)
{
    Slot *sl = slot_get(s);
    if (pyr_set_scales(sigma0, sigma_n, &s->gpyr) || pyr_set_scales(sigma0, sigma_n, &s->dog))
        return SIFT3D_FAILURE;
    if (!sl || !sl->have_image) return SIFT3D_SUCCESS;
    return gss_make(&s->gss, &s->gpyr);
}

int set_num_kp_levels_SIFT3D(SIFT3D *const sift3d, const unsigned int num_kp_levels)
{ /* sift.c:542-548 */
    return resize_all(sift3d, (int)num_kp_levels);
}

int set_sigma_n_SIFT3D(SIFT3D *const sift3d, const double sigma_n)
{ /* sift.c:552-564 */
    if (sigma_n < 0.0) {
        ERR("SIFT3D sigma_n must be nonnegative. Provided: %f \n", sigma_n);
        return SIFT3D_FAILURE;
    }
    return set_scales(sift3d, sift3d->gpyr.sigma0, sigma_n);
}

int set_sigma0_SIFT3D(SIFT3D *const sift3d, const double sigma0)
{ /* sift.c:568-580 */
    if (sigma0 < 0.0) {
        ERR("SIFT3D sigma0 must be nonnegative. Provided: %f \n", sigma0);
        return SIFT3D_FAILURE;
    }
    return set_scales(sift3d, sigma0, sift3d->gpyr.sigma_n);
}

int init_SIFT3D(SIFT3D *sift3d)
{ /* sift.c:583-626 */
    memset(sift3d, 0, sizeof(*sift3d));
    sift3d->gss.num_filters = -1;
    if (mesh_build(&sift3d->mesh)) return SIFT3D_FAILURE;
    if ((sift3d->kernels.downsample_2 = slot_new()) == 0) return SIFT3D_FAILURE;
    image_blank(&sift3d->im);
    sift3d->dog.first_level = sift3d->gpyr.first_level = -1;
    sift3d->dense_rotate = SIFT3D_FALSE;
    if (set_sigma_n_SIFT3D(sift3d, sigma_n_default) || set_sigma0_SIFT3D(sift3d, sigma0_default) ||
        set_peak_thresh_SIFT3D(sift3d, peak_thresh_default) ||
        set_corner_thresh_SIFT3D(sift3d, corner_thresh_default) ||
        set_num_kp_levels_SIFT3D(sift3d, num_kp_levels_default))
        return SIFT3D_FAILURE;
    return SIFT3D_SUCCESS;
}

void cleanup_SIFT3D(SIFT3D *const sift3d)
{ /* sift.c:659-678 */
    slot_release(sift3d);
    sift3d->kernels.downsample_2 = 0;
    free(sift3d->im.data);
    sift3d->im.data = NULL;
    pyr_free_host_data(&sift3d->gpyr, sift3d->gpyr.num_levels * sift3d->gpyr.num_octaves);
    pyr_free_host_data(&sift3d->dog, sift3d->dog.num_levels * sift3d->dog.num_octaves);
    free(sift3d->gpyr.levels);
    free(sift3d->dog.levels);
    sift3d->gpyr.levels = sift3d->dog.levels = NULL;
    gss_free(&sift3d->gss);
    free(sift3d->mesh.tri);
    sift3d->mesh.tri = NULL;
}

int copy_SIFT3D(const SIFT3D *const src, SIFT3D *const dst)
{ /* sift.c:629-655: deep copy, including the device-resident image and pyramids */
    Slot *ss, *ds;
    cleanup_SIFT3D(dst);
    if (init_SIFT3D(dst)) return SIFT3D_FAILURE;
    set_sigma_n_SIFT3D(dst, src->gpyr.sigma_n);
    set_sigma0_SIFT3D(dst, src->gpyr.sigma0);
    if (set_peak_thresh_SIFT3D(dst, src->peak_thresh) ||
        set_corner_thresh_SIFT3D(dst, src->corner_thresh) ||
        set_num_kp_levels_SIFT3D(dst, src->gpyr.num_kp_levels))
        return SIFT3D_FAILURE;
    dst->dense_rotate = src->dense_rotate;
    ss = slot_get(src);
    ds = slot_get(dst);
    if (ss && ds && ss->have_image && ss->eng) {
        s3d_engine *de = engine_of(dst);
        if (!de) return SIFT3D_FAILURE;
        dst->im = src->im;
        dst->im.data = NULL;
        ds->have_image = 1;
        if (resize_all(dst, src->gpyr.num_kp_levels)) return SIFT3D_FAILURE;
        memcpy(ds->resize_units, ss->resize_units, sizeof(ds->resize_units));
        if (push_geometry(dst, ds) || s3d_pyramid_copy(de, ss->eng)) return SIFT3D_FAILURE;
        /* the reference copies the levels' host data too (sift.c:650-651): do so when the
         * source has host copies */
        if (src->gpyr.levels && src->gpyr.num_levels * src->gpyr.num_octaves > 0 &&
            src->gpyr.levels[0].data && sift3d_b200_materialize_pyramids(dst))
            return SIFT3D_FAILURE;
    }
    return SIFT3D_SUCCESS;
}

void print_opts_SIFT3D(void)
{ /* sift.c:702-728 */
    printf("SIFT3D Options: \n"
           " --%s [value] \n"
           "    The smallest allowed absolute DoG value, as a fraction \n"
           "        of the largest. Must be on the interval (0, 1]. \n"
           "        (default: %.2f) \n"
           " --%s [value] \n"
           "    The smallest allowed corner score, on the interval \n"
           "        [0, 1]. (default: %.2f) \n"
           " --%s [value] \n"
           "    The number of pyramid levels per octave in which \n"
           "        keypoints are found. Must be a positive integer. \n"
           "        (default: %d) \n"
           " --%s [value] \n"
           "    The nominal scale parameter of the input data, on the \n"
           "        interval (0, inf). (default: %.2f) \n"
           " --%s [value] \n"
           "    The scale parameter of the first level of octave 0, on \n"
           "        the interval (0, inf). (default: %.2f) \n",
           opt_peak_thresh, peak_thresh_default, opt_corner_thresh, corner_thresh_default,
           opt_num_kp_levels, num_kp_levels_default, opt_sigma_n, sigma_n_default, opt_sigma0,
           sigma0_default);
}

int parse_args_SIFT3D(SIFT3D *const sift3d, const int argc, char **argv, const int check_err)
{ /* sift.c:754-879: consume the five SIFT3D long options, compact argv */
    enum { O_PEAK = 'a', O_CORNER, O_LEVELS, O_SIGMA_N, O_SIGMA0 };
    const struct option longopts[] = {{opt_peak_thresh, required_argument, NULL, O_PEAK},
                                      {opt_corner_thresh, required_argument, NULL, O_CORNER},
                                      {opt_num_kp_levels, required_argument, NULL, O_LEVELS},
                                      {opt_sigma_n, required_argument, NULL, O_SIGMA_N},
                                      {opt_sigma0, required_argument, NULL, O_SIGMA0},
                                      {0, 0, 0, 0}};
    const int opterr_start = opterr;
    unsigned char *used;
    int c, err = 0, i, kept = 0;
    opterr = check_err;
    if ((used = (unsigned char *)calloc(argc > 0 ? argc : 1, 1)) == NULL) {
        ERR("parse_args_SIFT3D: out of memory \n");
        return -1;
    }
    while ((c = getopt_long(argc, argv, "-", longopts, NULL)) != -1) {
        const int idx = optind - 1;
        const double dval = optarg ? atof(optarg) : 0.0;
        const int ival = optarg ? atoi(optarg) : 0;
        int rc = 0, mine = 1;
        switch (c) {
        case O_PEAK:
            rc = set_peak_thresh_SIFT3D(sift3d, dval);
            break;
        case O_CORNER:
            rc = set_corner_thresh_SIFT3D(sift3d, dval);
            break;
        case O_LEVELS:
            if (ival <= 0) {
                ERR("SIFT3D num_kp_levels must be positive. Provided: %d \n", ival);
                rc = -1;
            } else {
                rc = set_num_kp_levels_SIFT3D(sift3d, ival);
            }
            break;
        case O_SIGMA_N:
            set_sigma_n_SIFT3D(sift3d, dval);
            break;
        case O_SIGMA0:
            set_sigma0_SIFT3D(sift3d, dval);
            break;
        default: /* '?' and, because optstring is "-", every non-option argument (c == 1):
                  * the reference flags both when check_err is set (sift.c:844-848) */
            mine = 0;
            if (check_err) err = 1;
        }
        if (rc) {
            free(used);
            return -1;
        }
        if (mine && idx >= 1) used[idx - 1] = used[idx] = 1;
    }
    for (i = 0; i < argc; i++)
        if (!used[i]) argv[kept++] = argv[i];
    opterr = opterr_start;
    free(used);
    if (check_err && err) return -1;
    optind = 0;
    return kept;
}

/* ============================================================ hot path */

/* set_im_SIFT3D, sift.c:883-913: the copy goes straight to HBM; scaling happens
 * on the device at the head of s3d_build_pyramid. */
static int set_image(SIFT3D *const s, const Image *const im)
{
    Slot *sl = slot_get(s);
    s3d_engine *e = engine_of(s);
    const int first = !sl || !sl->have_image;
    const int changed = first || s->im.nx != im->nx || s->im.ny != im->ny || s->im.nz != im->nz;
    if (!e || !sl) return SIFT3D_FAILURE;
    if (im->data == NULL) return SIFT3D_FAILURE; /* im_copy_data, imutil.c:1901-1902 */
    if (im->nx < 1 || im->ny < 1 || im->nz < 1) {
        ERR("im_resize: invalid dimensions %d x %d x %d \n", im->nx, im->ny, im->nz);
        return SIFT3D_FAILURE;
    }
    if (s3d_image_upload(e, im->data, im->nx, im->ny, im->nz, im->xs, im->ys, im->zs))
        return SIFT3D_FAILURE;
    /* im_copy_dims (imutil.c:1873-1890): dims, strides, units, nc -- but not s */
    s->im.nx = im->nx, s->im.ny = im->ny, s->im.nz = im->nz;
    s->im.ux = im->ux, s->im.uy = im->uy, s->im.uz = im->uz;
    s->im.nc = im->nc;
    image_default_stride(&s->im);
    s->im.size = (size_t)im->nx * im->ny * im->nz * im->nc;
    sl->have_image = 1;
    if (changed && resize_all(s, s->gpyr.num_kp_levels)) return SIFT3D_FAILURE;
    return push_geometry(s, sl);
}

/* Z-slab tiling: `im` holds this rank's planes [zsplit[rank], zsplit[rank+1]) of a volume of
 * zsplit[nranks] planes; sift3d->im takes the GLOBAL dims so that the pyramid geometry
 * (resize_SIFT3D, sift.c:938-986) and verify_keys (sift.c:2050) see the whole volume. */
static int set_image_slab(SIFT3D *const s, const Image *const im, const int *zsplit, s3d_comm *comm)
{
    Slot *sl = slot_get(s);
    s3d_engine *e = engine_of(s);
    const int rank = s3d_comm_rank(comm), nr = s3d_comm_size(comm);
    const int NZ = zsplit[nr];
    if (!e || !sl) return SIFT3D_FAILURE;
    if (im->nz != zsplit[rank + 1] - zsplit[rank] || (im->nz > 0 && im->data == NULL) ||
        im->nx < 1 || im->ny < 1 || NZ < 1) {
        ERR("SIFT3D_detect_keypoints_slab: slab of %d planes does not match the split [%d, %d) \n",
            im->nz, zsplit[rank], zsplit[rank + 1]);
        return SIFT3D_FAILURE;
    }
    s->im.nx = im->nx, s->im.ny = im->ny, s->im.nz = NZ;
    s->im.ux = im->ux, s->im.uy = im->uy, s->im.uz = im->uz;
    s->im.nc = im->nc;
    image_default_stride(&s->im);
    s->im.size = (size_t)im->nx * im->ny * NZ * im->nc;
    sl->have_image = 1;
    if (resize_all(s, s->gpyr.num_kp_levels)) return SIFT3D_FAILURE;
    if (push_geometry_ex(s, sl, zsplit, comm)) return SIFT3D_FAILURE;
    return s3d_slab_image_upload(e, im->data, im->xs, im->ys, im->zs) ? SIFT3D_FAILURE
                                                                      : SIFT3D_SUCCESS;
}

static int detect_common(SIFT3D *const sift3d, const Image *const im, Keypoint_store *const kp,
                         const int *zsplit, s3d_comm *comm);

int SIFT3D_detect_keypoints(SIFT3D *const sift3d, const Image *const im, Keypoint_store *const kp)
{ /* sift.c:1609-1641 */
    return detect_common(sift3d, im, kp, NULL, NULL);
}

/* Extension (no reference counterpart): SIFT3D_detect_keypoints on one Z-slab of a volume
 * that is tiled over the ranks of `comm` (s3d_comm_create_nccl / _local).  `kp` receives this
 * rank's keypoints -- global coordinates, reference scan order -- so that concatenating the
 * ranks' lists per (octave, level) reproduces the single-volume result.  A following
 * SIFT3D_extract_descriptors works on them unchanged. */
int SIFT3D_detect_keypoints_slab(SIFT3D *const sift3d, const Image *const im, const int *zsplit,
                                 void *comm, Keypoint_store *const kp)
{
    if (!comm || !zsplit) return SIFT3D_FAILURE;
    return detect_common(sift3d, im, kp, zsplit, (s3d_comm *)comm);
}

static int detect_common(SIFT3D *const sift3d, const Image *const im, Keypoint_store *const kp,
                         const int *zsplit, s3d_comm *comm)
{
    Slot *sl;
    s3d_engine *e;
    s3d_keypoint *tmp = NULL;
    int ncand = 0, nkp = 0, i;
    if (im->nc != 1) {
        ERR("SIFT3D_detect_keypoints: invalid number of image channels: %d -- only "
            "single-channel images are supported \n", im->nc);
        return SIFT3D_FAILURE;
    }
    /* host copies of the previous pyramid (sift3d_b200_materialize_pyramids) are stale now */
    pyr_free_host_data(&sift3d->gpyr, sift3d->gpyr.levels ? sift3d->gpyr.num_levels * sift3d->gpyr.num_octaves : 0);
    pyr_free_host_data(&sift3d->dog, sift3d->dog.levels ? sift3d->dog.num_levels * sift3d->dog.num_octaves : 0);
    if (comm ? set_image_slab(sift3d, im, zsplit, comm) : set_image(sift3d, im))
        return SIFT3D_FAILURE;
    sl = slot_get(sift3d);
    e = sl->eng;
    if (sift3d->dog.num_levels < 3) { /* detect_extrema, sift.c:1089-1093 */
        printf("detect_extrema: Requires at least 3 levels per octave, provided only %d \n",
               sift3d->dog.num_levels);
        return SIFT3D_FAILURE;
    }
    if (s3d_build_pyramid(e) || s3d_detect_extrema(e, sift3d->peak_thresh, &ncand) ||
        s3d_assign_orientations(e, sift3d->corner_thresh, &nkp))
        return SIFT3D_FAILURE;
    sl->ncand = ncand;
    {   /* kp dims = DoG level (0, 0) (sift.c:1096-1099) */
        const Image *l = PYR_LEVEL(&sift3d->dog, 0, 0);
        kp->nx = l->nx, kp->ny = l->ny, kp->nz = l->nz;
    }
    if (!comm && getenv("SIFT3D_HOST_PYRAMID") && atoi(getenv("SIFT3D_HOST_PYRAMID")) &&
        sift3d_b200_materialize_pyramids(sift3d))
        return SIFT3D_FAILURE;
    if (resize_Keypoint_store(kp, (size_t)nkp)) return SIFT3D_FAILURE;
    if (nkp == 0) return SIFT3D_SUCCESS;
    if ((tmp = (s3d_keypoint *)malloc((size_t)nkp * sizeof(s3d_keypoint))) == NULL)
        return SIFT3D_FAILURE;
    if (s3d_keypoints_download(e, tmp, nkp)) {
        free(tmp);
        return SIFT3D_FAILURE;
    }
    for (i = 0; i < nkp; i++) {
        Keypoint *k = kp->buf + i;
        init_Keypoint(k);
        memcpy(k->r_data, tmp[i].R, sizeof(k->r_data));
        k->xd = (double)tmp[i].x;
        k->yd = (double)tmp[i].y;
        k->zd = (double)tmp[i].z;
        k->sd = tmp[i].sd;
        k->o = tmp[i].o;
        k->s = tmp[i].s;
    }
    free(tmp);
    return SIFT3D_SUCCESS;
}

int SIFT3D_have_gpyr(const SIFT3D *const sift3d)
{ /* sift.c:1936-1942 */
    const Pyramid *const g = &sift3d->gpyr;
    return g->levels != NULL && g->num_levels != 0 && g->num_octaves != 0;
}

/* verify_keys, sift.c:2050-2091 */
static int verify_keys(const Keypoint_store *const kp, const Image *const im)
{
    const int num = (int)kp->slab.num;
    int i;
    if (num < 1) {
        ERR("verify_keys: invalid number of keypoints: %d \n", num);
        return SIFT3D_FAILURE;
    }
    for (i = 0; i < num; i++) {
        const Keypoint *key = kp->buf + i;
        const double f = ldexp(1.0, key->o);
        if (key->xd < 0 || key->yd < 0 || key->zd < 0 || key->xd * f >= (double)im->nx ||
            key->yd * f >= (double)im->ny || key->zd * f >= (double)im->nz) {
            ERR("verify_keys: keypoint %d (%f, %f, %f) octave %d exceeds image dimensions "
                "(%d, %d, %d) \n", i, key->xd, key->yd, key->zd, key->o, im->nx, im->ny, im->nz);
            return SIFT3D_FAILURE;
        }
        if (key->sd <= 0) {
            ERR("verify_keys: keypoint %d has invalid scale %f \n", i, key->sd);
            return SIFT3D_FAILURE;
        }
    }
    return SIFT3D_SUCCESS;
}

static s3d_keypoint *pack_keys(const Keypoint_store *kp, int to_base)
{
    const int n = (int)kp->slab.num;
    s3d_keypoint *out = (s3d_keypoint *)malloc((size_t)n * sizeof(s3d_keypoint));
    int i, j;
    if (!out) return NULL;
    for (i = 0; i < n; i++) {
        const Keypoint *k = kp->buf + i;
        const double f = to_base ? ldexp(1.0, k->o) : 1.0; /* keypoint2base, sift.c:2094-2115 */
        for (j = 0; j < 9; j++) out[i].R[j] = k->R.u.data_float ? k->R.u.data_float[j] : 0.0f;
        out[i].x = (float)(k->xd * f);
        out[i].y = (float)(k->yd * f);
        out[i].z = (float)(k->zd * f);
        out[i].sd = k->sd;
        out[i].o = to_base ? 0 : k->o;
        out[i].s = to_base ? 0 : k->s;
    }
    return out;
}

int SIFT3D_extract_descriptors(SIFT3D *const sift3d, const Keypoint_store *const kp,
                               SIFT3D_Descriptor_store *const desc)
{ /* sift.c:2025-2046 + _SIFT3D_extract_descriptors sift.c:2207-2243 */
    const Image *first;
    s3d_engine *e;
    s3d_keypoint *keys;
    int rc;
    if (verify_keys(kp, &sift3d->im)) return SIFT3D_FAILURE;
    if (!SIFT3D_have_gpyr(sift3d)) {
        ERR("SIFT3D_extract_descriptors: no Gaussian pyramid is available. Make sure "
            "SIFT3D_detect_keypoints was called prior to calling this function. \n");
        return SIFT3D_FAILURE;
    }
    if ((e = engine_of(sift3d)) == NULL) return SIFT3D_FAILURE;
    first = PYR_LEVEL(&sift3d->gpyr, sift3d->gpyr.first_octave, sift3d->gpyr.first_level);
    desc->nx = first->nx, desc->ny = first->ny, desc->nz = first->nz;
    if (desc_store_resize(desc, (int)kp->slab.num)) return SIFT3D_FAILURE;
    if ((keys = pack_keys(kp, 0)) == NULL) return SIFT3D_FAILURE;
    rc = s3d_extract_descriptors(e, keys, (int)kp->slab.num, desc->buf);
    free(keys);
    return rc ? SIFT3D_FAILURE : SIFT3D_SUCCESS;
}

/* raw-image helpers: a private one-level engine (the reference builds a local
 * Pyramid and leaves sift3d->gpyr alone, sift.c:2135-2165) */
static s3d_engine *raw_engine(const SIFT3D *s, const Image *im)
{
    s3d_engine *e = NULL, *parent = engine_of(s);
    Gauss_filter g;
    s3d_filter f;
    const double units[3] = {im->ux, im->uy, im->uz};
    float v[ICOS_NFACES * 9];
    int idx[ICOS_NFACES * 3], i, j, rc;
    if (!parent) return NULL;
    if (im->nc != 1) {
        ERR("sift3d_b200: raw-image calls support single-channel images only \n");
        return NULL;
    }
    if (s3dh_gauss_incremental(&g, s->gpyr.sigma_n, s->gpyr.sigma0, 3)) return NULL;
    if (s3d_engine_create(&e, s3d_engine_device(parent))) {
        free(g.f.kernel);
        return NULL;
    }
    for (i = 0; i < ICOS_NFACES; i++)
        for (j = 0; j < 3; j++) {
            v[9 * i + 3 * j + 0] = s->mesh.tri[i].v[j].x;
            v[9 * i + 3 * j + 1] = s->mesh.tri[i].v[j].y;
            v[9 * i + 3 * j + 2] = s->mesh.tri[i].v[j].z;
            idx[3 * i + j] = s->mesh.tri[i].idx[j];
        }
    f.taps = g.f.kernel;
    f.width = g.f.width;
    rc = s3d_set_mesh(e, v, idx) ||
         s3d_single_level(e, im->data, im->nx, im->ny, im->nz, im->xs, im->ys, im->zs, units,
                          s->gpyr.sigma0, &f);
    free(g.f.kernel);
    if (rc) {
        s3d_engine_destroy(e);
        return NULL;
    }
    return e;
}

int SIFT3D_extract_raw_descriptors(SIFT3D *const sift3d, const Image *const im,
                                   const Keypoint_store *const kp,
                                   SIFT3D_Descriptor_store *const desc)
{ /* sift.c:2131-2195 */
    s3d_engine *e;
    s3d_keypoint *keys;
    int rc;
    if (verify_keys(kp, im)) return SIFT3D_FAILURE;
    if ((e = raw_engine(sift3d, im)) == NULL) return SIFT3D_FAILURE;
    desc->nx = im->nx, desc->ny = im->ny, desc->nz = im->nz;
    if (desc_store_resize(desc, (int)kp->slab.num) || (keys = pack_keys(kp, 1)) == NULL) {
        s3d_engine_destroy(e);
        return SIFT3D_FAILURE;
    }
    rc = s3d_extract_descriptors(e, keys, (int)kp->slab.num, desc->buf);
    free(keys);
    s3d_engine_destroy(e);
    return rc ? SIFT3D_FAILURE : SIFT3D_SUCCESS;
}

int SIFT3D_assign_orientations(const SIFT3D *const sift3d, const Image *const im,
                               Keypoint_store *const kp, double **const conf)
{ /* sift.c:1534-1604: sigma = key_base.sd, no corner threshold; REJECT -> identity, conf -1 */
    const int num = (int)kp->slab.num;
    s3d_engine *e;
    s3d_keypoint *keys;
    unsigned char *ok;
    int i, rc;
    if (verify_keys(kp, im)) return SIFT3D_FAILURE;
    if ((*conf = (double *)safe_realloc(*conf, num * sizeof(double))) == NULL)
        return SIFT3D_FAILURE;
    if ((e = raw_engine(sift3d, im)) == NULL) return SIFT3D_FAILURE;
    keys = pack_keys(kp, 1);
    ok = (unsigned char *)malloc(num);
    rc = (!keys || !ok) ? -1 : s3d_orient_keypoints(e, keys, num, 1.0, -1.0, *conf, ok);
    for (i = 0; !rc && i < num; i++) {
        Keypoint *k = kp->buf + i;
        init_Keypoint(k);
        if (ok[i]) {
            memcpy(k->r_data, keys[i].R, sizeof(k->r_data));
        } else {
            const float I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
            memcpy(k->r_data, I3, sizeof(I3));
            (*conf)[i] = -1.0;
        }
    }
    free(keys);
    free(ok);
    s3d_engine_destroy(e);
    return rc ? SIFT3D_FAILURE : SIFT3D_SUCCESS;
}

int SIFT3D_extract_dense_descriptors(SIFT3D *const sift3d, const Image *const in, Image *const desc)
{ /* sift.c:2354-2424 (dense_rotate == 0: sift.c:2429-2496) */
    const double units[3] = {in->ux, in->uy, in->uz};
    const double desc_units[3] = {desc->ux, desc->uy, desc->uz};
    const double sigma_win = sift3d->gpyr.sigma0 * desc_sig_fctr / NHIST_PER_DIM;
    Gauss_filter gs, gw;
    s3d_filter fs, fw;
    s3d_engine *e;
    size_t size;
    int rc;
    if (in->nc != 1) {
        ERR("SIFT3D_extract_dense_descriptors: invalid number of channels: %d. This function "
            "only supports single-channel images. \n", in->nc);
        return SIFT3D_FAILURE;
    }
    if (in->data == NULL || in->nx < 1 || in->ny < 1 || in->nz < 1) return SIFT3D_FAILURE;
    /* resize the output: dims of `in`, 12 channels, default stride; units untouched
     * (sift.c:2375-2380) */
    desc->nx = in->nx, desc->ny = in->ny, desc->nz = in->nz;
    desc->nc = HIST_NUMEL;
    image_default_stride(desc);
    size = (size_t)desc->nx * desc->ny * desc->nz * desc->nc;
    if (desc->size != size) {
        desc->size = size;
        if ((desc->data = (float *)safe_realloc(desc->data, size * sizeof(float))) == NULL) {
            desc->size = 0;
            return SIFT3D_FAILURE;
        }
    }
    if ((e = engine_of(sift3d)) == NULL) return SIFT3D_FAILURE;
    if (s3dh_gauss_incremental(&gs, sift3d->gpyr.sigma_n, sift3d->gpyr.sigma0, 3))
        return SIFT3D_FAILURE;
    if (sift3d->dense_rotate) { /* sift.c:2521-2588 */
        fs.taps = gs.f.kernel, fs.width = gs.f.width;
        rc = s3d_dense_descriptors_rotate(e, in->data, in->nx, in->ny, in->nz, in->xs, in->ys,
                                          in->zs, units, &fs, sift3d->gpyr.sigma0 * ori_sig_fctr,
                                          sigma_win, sift3d->corner_thresh, desc->data);
        free(gs.f.kernel);
        return rc ? SIFT3D_FAILURE : SIFT3D_SUCCESS;
    }
    if (s3dh_gauss_filter(&gw, sigma_win, 3)) {
        free(gs.f.kernel);
        return SIFT3D_FAILURE;
    }
    fs.taps = gs.f.kernel, fs.width = gs.f.width;
    fw.taps = gw.f.kernel, fw.width = gw.f.width;
    rc = s3d_dense_descriptors(e, in->data, in->nx, in->ny, in->nz, in->xs, in->ys, in->zs, units,
                               desc_units, &fs, &fw, desc->data);
    free(gs.f.kernel);
    free(gw.f.kernel);
    return rc ? SIFT3D_FAILURE : SIFT3D_SUCCESS;
}

int sift3d_b200_fetch_level(const SIFT3D *sift3d, int which, int o, int s, float *dst)
{
    s3d_engine *e = engine_of(sift3d);
    return (!e || s3d_level_download(e, which, o, s, dst)) ? SIFT3D_FAILURE : SIFT3D_SUCCESS;
}

/* Extension: host copies of every Gaussian and DoG level, so that code which reads
 * sift3d->gpyr.levels[i].data / sift3d->dog.levels[i].data as it would after the reference's
 * SIFT3D_detect_keypoints (write_pyramid, imutil.c:4093; copy_SIFT3D's deep copy,
 * sift.c:650-651) finds the same bits there.  The data is malloc memory owned by the SIFT3D
 * object (freed by cleanup_SIFT3D or the next resize) and is valid until the next detect
 * call.  With $SIFT3D_HOST_PYRAMID=1 every SIFT3D_detect_keypoints ends with this call
 * (6.3 GB of D2H for a 512^3 volume: opt-in). */
int sift3d_b200_materialize_pyramids(SIFT3D *sift3d)
{
    s3d_engine *e = engine_of(sift3d);
    int which, o, s;
    if (!e || !SIFT3D_have_gpyr(sift3d)) return SIFT3D_FAILURE;
    for (which = 0; which < 2; which++) {
        Pyramid *pyr = which ? &sift3d->dog : &sift3d->gpyr;
        for (o = pyr->first_octave; o < pyr->first_octave + pyr->num_octaves; o++)
            for (s = pyr->first_level; s < pyr->first_level + pyr->num_levels; s++) {
                Image *l = PYR_LEVEL(pyr, o, s);
                const size_t n = (size_t)l->nx * l->ny * l->nz;
                if (n == 0) continue;
                if (l->size < n) {
                    float *d = (float *)safe_realloc(l->data, n * sizeof(float));
                    if (!d) return SIFT3D_FAILURE;
                    l->data = d;
                    l->size = n;
                }
                if (s3d_level_download(e, which, o, s, l->data)) return SIFT3D_FAILURE;
            }
    }
    return SIFT3D_SUCCESS;
}

/* ============================================================ converters (host) */

int Keypoint_store_to_Mat_rm(const Keypoint_store *const kp, Mat_rm *const mat)
{ /* sift.c:2597-2624 */
    const int num = (int)kp->slab.num;
    int i;
    mat->num_rows = num;
    mat->num_cols = IM_NDIMS;
    mat->type = SIFT3D_DOUBLE;
    if (mat_resize(mat)) return SIFT3D_FAILURE;
    for (i = 0; i < num; i++) {
        const Keypoint *const key = kp->buf + i;
        const double f = ldexp(1.0, key->o);
        mat->u.data_double[3 * i + 0] = f * key->xd;
        mat->u.data_double[3 * i + 1] = f * key->yd;
        mat->u.data_double[3 * i + 2] = f * key->zd;
    }
    return SIFT3D_SUCCESS;
}

int SIFT3D_Descriptor_coords_to_Mat_rm(const SIFT3D_Descriptor_store *const store,
                                       Mat_rm *const mat)
{ /* sift.c:2628-2662 */
    const int n = (int)store->num;
    int i;
    if (n < 1) {
        printf("SIFT3D_Descriptor_coords_to_Mat_rm: invalid number of descriptors: %d \n", n);
        return SIFT3D_FAILURE;
    }
    mat->type = SIFT3D_DOUBLE;
    mat->num_rows = n;
    mat->num_cols = IM_NDIMS;
    if (mat_resize(mat)) return SIFT3D_FAILURE;
    for (i = 0; i < n; i++) {
        mat->u.data_double[3 * i + 0] = store->buf[i].xd;
        mat->u.data_double[3 * i + 1] = store->buf[i].yd;
        mat->u.data_double[3 * i + 2] = store->buf[i].zd;
    }
    return SIFT3D_SUCCESS;
}

int SIFT3D_Descriptor_store_to_Mat_rm(const SIFT3D_Descriptor_store *const store,
                                      Mat_rm *const mat)
{ /* sift.c:2674-2717: rows of [x y z el0 .. el767], float */
    const int n = (int)store->num, cols = IM_NDIMS + DESC_NUMEL;
    int i;
    if (n < 1) {
        printf("SIFT3D_Descriptor_store_to_Mat_rm: invalid number of descriptors: %d \n", n);
        return SIFT3D_FAILURE;
    }
    mat->type = SIFT3D_FLOAT;
    mat->num_rows = n;
    mat->num_cols = cols;
    if (mat_resize(mat)) return SIFT3D_FAILURE;
    for (i = 0; i < n; i++) {
        float *row = mat->u.data_float + (size_t)i * cols;
        row[0] = (float)store->buf[i].xd;
        row[1] = (float)store->buf[i].yd;
        row[2] = (float)store->buf[i].zd;
        memcpy(row + IM_NDIMS, store->buf[i].hists, DESC_NUMEL * sizeof(float));
    }
    return SIFT3D_SUCCESS;
}

int Mat_rm_to_SIFT3D_Descriptor_store(const Mat_rm *const mat,
                                      SIFT3D_Descriptor_store *const store)
{ /* sift.c:2721-2768 */
    const int n = mat->num_rows, cols = mat->num_cols;
    int i;
    if (n < 1 || cols != IM_NDIMS + DESC_NUMEL) {
        ERR("Mat_rm_to_SIFT3D_Descriptor_store: invalid matrix dimensions: [%d X %d] \n", n, cols);
        return SIFT3D_FAILURE;
    }
    if (mat->type != SIFT3D_FLOAT) {
        ERR("Mat_rm_to_SIFT3D_Descriptor_store: matrix must have type SIFT3D_FLOAT");
        return SIFT3D_FAILURE;
    }
    if (desc_store_resize(store, n)) return SIFT3D_FAILURE;
    for (i = 0; i < n; i++) {
        const float *row = mat->u.data_float + (size_t)i * cols;
        store->buf[i].xd = row[0];
        store->buf[i].yd = row[1];
        store->buf[i].zd = row[2];
        store->buf[i].sd = sigma0_default;
        memcpy(store->buf[i].hists, row + IM_NDIMS, DESC_NUMEL * sizeof(float));
    }
    return SIFT3D_SUCCESS;
}

int SIFT3D_matches_to_Mat_rm(SIFT3D_Descriptor_store *d1, SIFT3D_Descriptor_store *d2,
                             const int *const matches, Mat_rm *const match1, Mat_rm *const match2)
{ /* sift.c:2784-2826 */
    const int num = (int)d1->num;
    int i, m = 0;
    match1->num_rows = match2->num_rows = num;
    match1->num_cols = match2->num_cols = 3;
    match1->type = match2->type = SIFT3D_DOUBLE;
    if (mat_resize(match1) || mat_resize(match2)) return SIFT3D_FAILURE;
    for (i = 0; i < num; i++) {
        if (matches[i] == -1) continue;
        match1->u.data_double[3 * m + 0] = d1->buf[i].xd;
        match1->u.data_double[3 * m + 1] = d1->buf[i].yd;
        match1->u.data_double[3 * m + 2] = d1->buf[i].zd;
        match2->u.data_double[3 * m + 0] = d2->buf[matches[i]].xd;
        match2->u.data_double[3 * m + 1] = d2->buf[matches[i]].yd;
        match2->u.data_double[3 * m + 2] = d2->buf[matches[i]].zd;
        m++;
    }
    match1->num_rows = match2->num_rows = m;
    if (mat_resize(match1) || mat_resize(match2)) return SIFT3D_FAILURE;
    return SIFT3D_SUCCESS;
}

/* match_desc, sift.c:2892-2969.  The reference's early exit never changes the
 * outcome (a partial sum above ssd_nearest can only grow), so the plain full sum
 * below returns the same index. */
/* Engine for the entry points that take no SIFT3D object (matching, resampling). */
static s3d_engine *g_util_engine = NULL;

static s3d_engine *util_engine(void)
{
    s3d_engine *e;
    pthread_mutex_lock(&g_lock);
    if (!g_util_engine && s3d_engine_create(&g_util_engine, -1)) {
        ERR("sift3d_b200: cannot create the CUDA engine: %s \n", s3d_last_create_error());
        g_util_engine = NULL;
    }
    e = g_util_engine;
    pthread_mutex_unlock(&g_lock);
    return e;
}

int SIFT3D_nn_match(const SIFT3D_Descriptor_store *const d1,
                    const SIFT3D_Descriptor_store *const d2, const float nn_thresh,
                    int **const matches)
{ /* sift.c:2840-2888: forward-backward consistent nearest neighbour with ratio test; the
   * exhaustive f64 SSD search (match_desc, sift.c:2893-2969) runs on the device */
    const int num = (int)d1->num;
    s3d_engine *e;
    int i, rc;
    if (num < 1) {
        ERR("_SIFT3D_nn_match: invalid number of descriptors in d1: %d \n", num);
        return SIFT3D_FAILURE;
    }
    if ((*matches = (int *)safe_realloc(*matches, num * sizeof(int))) == NULL) {
        ERR("_SIFT3D_nn_match: out of memory! \n");
        return SIFT3D_FAILURE;
    }
    for (i = 0; i < num; i++) (*matches)[i] = -1;
    if (d2->num < 1) return SIFT3D_SUCCESS; /* match_desc finds nothing in an empty store */
    if ((e = util_engine()) == NULL) return SIFT3D_FAILURE; /* no CPU fallback */
    pthread_mutex_lock(&g_util_lock);
    rc = s3d_nn_match(e, d1->buf, num, d2->buf, (int)d2->num, nn_thresh, *matches);
    pthread_mutex_unlock(&g_util_lock);
    return rc ? SIFT3D_FAILURE : SIFT3D_SUCCESS;
}

/* ============================================================ resampling
 * The reference keeps these in libimutil (im_inv_transform imutil.c:2040, im_resample
 * imutil.c:2191); they are exported under sift3d_b200_* names so that both libraries can be
 * loaded side by side -- INTEGRATION.md shows the two-line forwarding a maintainer adds. */

static int image_resize_data(Image *im)
{ /* im_resize, imutil.c:1523-1560: realloc when the element count changes */
    const size_t size = (size_t)im->nx * im->ny * im->nz * im->nc;
    if (im->nx < 1 || im->ny < 1 || im->nz < 1 || im->nc < 1) {
        ERR("im_resize: invalid dimensions %d x %d x %d x %d \n", im->nx, im->ny, im->nz, im->nc);
        return SIFT3D_FAILURE;
    }
    if (im->size != size || im->data == NULL) {
        im->size = size;
        if ((im->data = (float *)safe_realloc(im->data, size * sizeof(float))) == NULL) {
            im->size = 0;
            return SIFT3D_FAILURE;
        }
    }
    return SIFT3D_SUCCESS;
}

/* im_inv_transform for an affine map given as a row-major 3x4 matrix (the `A` of an Affine,
 * imtypes.h:374-377): dst voxel (x,y,z) takes the value of src at A*[x y z 1].
 * interp: 0 = LINEAR, 1 = LANCZOS2 (interp_type, imtypes.h:343-346). */
int sift3d_b200_im_inv_transform_affine(const double A[12], const Image *const src, const int interp,
                                        const int resize, Image *const dst)
{
    s3d_engine *e;
    float *tmp = NULL;
    const float *data = src->data;
    int rc;
    if (src->data == NULL) return SIFT3D_FAILURE;
    if (interp != 0 && interp != 1) {
        ERR("im_inv_transform: unrecognized interpolation type");
        return SIFT3D_FAILURE;
    }
    if (resize) { /* im_copy_dims, imutil.c:1873-1890 */
        dst->nx = src->nx, dst->ny = src->ny, dst->nz = src->nz, dst->nc = src->nc;
        dst->ux = src->ux, dst->uy = src->uy, dst->uz = src->uz;
        image_default_stride(dst);
        if (image_resize_data(dst)) return SIFT3D_FAILURE;
    }
    if (dst->data == NULL || dst->nc != src->nc) return SIFT3D_FAILURE;
    if ((e = util_engine()) == NULL) return SIFT3D_FAILURE;
    if (src->xs != (size_t)src->nc || src->ys != (size_t)src->nc * src->nx ||
        src->zs != (size_t)src->nc * src->nx * src->ny) { /* custom strides: gather first */
        int x, y, z, c;
        if ((tmp = (float *)malloc((size_t)src->nx * src->ny * src->nz * src->nc * sizeof(float))) == NULL)
            return SIFT3D_FAILURE;
        for (z = 0; z < src->nz; z++)
            for (y = 0; y < src->ny; y++)
                for (x = 0; x < src->nx; x++)
                    for (c = 0; c < src->nc; c++)
                        tmp[c + (size_t)src->nc * (x + (size_t)src->nx * (y + (size_t)src->ny * z))] =
                            src->data[c + x * src->xs + y * src->ys + z * src->zs];
        data = tmp;
    }
    {   /* the reference writes through SIFT3D_IM_GET_VOX, i.e. the caller's dst strides
         * (imutil.c:2062-2076): resample into a contiguous buffer and scatter when they are not
         * the default ones; dst->xs/ys/zs are never modified */
        const int dst_default = dst->xs == (size_t)dst->nc && dst->ys == (size_t)dst->nc * dst->nx &&
                                dst->zs == (size_t)dst->nc * dst->nx * dst->ny;
        float *out = dst->data, *otmp = NULL;
        const size_t dn = (size_t)dst->nx * dst->ny * dst->nz * dst->nc;
        if (!dst_default) {
            if ((otmp = (float *)malloc(dn * sizeof(float))) == NULL) {
                free(tmp);
                return SIFT3D_FAILURE;
            }
            out = otmp;
        }
        pthread_mutex_lock(&g_util_lock);
        rc = s3d_resample_affine(e, data, src->nx, src->ny, src->nz, src->nc, A, interp, out,
                                 dst->nx, dst->ny, dst->nz);
        pthread_mutex_unlock(&g_util_lock);
        if (otmp && !rc) {
            int x, y, z, c;
            for (z = 0; z < dst->nz; z++)
                for (y = 0; y < dst->ny; y++)
                    for (x = 0; x < dst->nx; x++)
                        for (c = 0; c < dst->nc; c++)
                            dst->data[c + x * dst->xs + y * dst->ys + z * dst->zs] =
                                otmp[c + (size_t)dst->nc * (x + (size_t)dst->nx * (y + (size_t)dst->ny * z))];
        }
        free(otmp);
    }
    free(tmp);
    return rc ? SIFT3D_FAILURE : SIFT3D_SUCCESS;
}

/* im_resample, imutil.c:2191-2244: resample to new voxel units. */
int sift3d_b200_im_resample(const Image *const src, const double *const units, const int interp,
                            Image *const dst)
{
    double A[12] = {0}, factors[3];
    const double su[3] = {src->ux, src->uy, src->uz};
    const int sd[3] = {src->nx, src->ny, src->nz};
    int dims[3], i;
    for (i = 0; i < 3; i++) {
        factors[i] = su[i] / units[i];
        A[4 * i + i] = 1.0 / factors[i];
        dims[i] = (int)ceil((double)sd[i] * factors[i]);
    }
    dst->nc = src->nc;
    dst->nx = dims[0], dst->ny = dims[1], dst->nz = dims[2];
    image_default_stride(dst);
    if (image_resize_data(dst)) return SIFT3D_FAILURE;
    if (sift3d_b200_im_inv_transform_affine(A, src, interp, 0, dst)) return SIFT3D_FAILURE;
    dst->ux = units[0], dst->uy = units[1], dst->uz = units[2];
    return SIFT3D_SUCCESS;
}

/* ---- functions that lean on libimutil when it is linked (drop-in deployments) */
extern int init_im_with_dims(Image *const, const int, const int, const int, const int)
    __attribute__((weak));
extern int im_pad(const Image *const, Image *const) __attribute__((weak));
extern int im_concat(const Image *const, const Image *const, const int, Image *const)
    __attribute__((weak));
extern int draw_points(const Mat_rm *const, const int *const, const int, Image *const)
    __attribute__((weak));
extern int draw_lines(const Mat_rm *const, const Mat_rm *const, const int *const, Image *const)
    __attribute__((weak));

static int write_csv(const char *path, const Mat_rm *mat)
{ /* write_Mat_rm's text format (imutil.c:1343-1421), bytes identical to its "%f" / "%d" fields;
   * formatted in parallel by csv_io.c -- always ours, also when libimutil is linked */
    return sift3d_b200_write_Mat_rm(path, mat);
}

int write_Keypoint_store(const char *path, const Keypoint_store *const kp)
{ /* sift.c:3143-3202: x y z o s R00..R22 per row */
    const int rows = (int)kp->slab.num, cols = 5 + 9;
    Mat_rm mat;
    int i, j, rc;
    mat_init_empty(&mat, SIFT3D_DOUBLE);
    mat.num_rows = rows;
    mat.num_cols = cols;
    if (mat_resize(&mat)) return SIFT3D_FAILURE;
    for (i = 0; i < rows; i++) {
        const Keypoint *k = kp->buf + i;
        double *row = mat.u.data_double + (size_t)i * cols;
        row[0] = k->xd, row[1] = k->yd, row[2] = k->zd, row[3] = k->o, row[4] = k->sd;
        for (j = 0; j < 9; j++) row[5 + j] = (double)k->R.u.data_float[j];
    }
    rc = write_csv(path, &mat);
    free(mat.u.data_double);
    return rc;
}

int write_SIFT3D_Descriptor_store(const char *path, const SIFT3D_Descriptor_store *const desc)
{ /* sift.c:3206-3230 */
    Mat_rm mat;
    int rc;
    mat_init_empty(&mat, SIFT3D_FLOAT);
    if (SIFT3D_Descriptor_store_to_Mat_rm(desc, &mat)) {
        free(mat.u.data_double);
        return SIFT3D_FAILURE;
    }
    rc = write_csv(path, &mat);
    free(mat.u.data_double);
    return rc;
}

int draw_matches(const Image *const left, const Image *const right, const Mat_rm *const keys_left,
                 const Mat_rm *const keys_right, const Mat_rm *const match_left,
                 const Mat_rm *const match_right, Image *const concat, Image *const keys,
                 Image *const lines)
{ /* sift.c:2990-3128 -- visualisation, outside the accelerated path; needs libimutil */
    Image ctmp, lp, rp;
    Mat_rm kd, md;
    Image *carg = concat ? concat : &ctmp;
    const double right_pad = (double)left->nx;
    const int ny_pad = MAXV(right->ny, left->ny), nz_pad = MAXV(right->nz, left->nz);
    int rc = SIFT3D_FAILURE, i, j;
    if (!init_im_with_dims || !im_pad || !im_concat || !draw_points || !draw_lines) {
        ERR("draw_matches: this build delegates drawing to libimutil, which is not linked \n");
        return SIFT3D_FAILURE;
    }
    if (!concat && !keys && !lines) {
        ERR("draw_matches: all outputs are NULL \n");
        return SIFT3D_FAILURE;
    }
    if ((keys && (!keys_left || !keys_right)) || (lines && (!match_left || !match_right))) {
        ERR("draw_matches: missing keypoint/match inputs for the requested outputs \n");
        return SIFT3D_FAILURE;
    }
    image_blank(&ctmp);
    image_blank(&lp);
    image_blank(&rp);
    mat_init_empty(&kd, SIFT3D_DOUBLE);
    mat_init_empty(&md, SIFT3D_DOUBLE);
    if (init_im_with_dims(&rp, right->nx, ny_pad, nz_pad, 1) ||
        init_im_with_dims(&lp, left->nx, ny_pad, nz_pad, 1) || im_pad(right, &rp) ||
        im_pad(left, &lp) || im_concat(&lp, &rp, 0, carg))
        goto done;
#define MAT_AT(m, r, c)                                                              \
    ((m)->type == SIFT3D_DOUBLE ? (m)->u.data_double[(size_t)(r) * (m)->num_cols + (c)] \
     : (m)->type == SIFT3D_FLOAT ? (double)(m)->u.data_float[(size_t)(r) * (m)->num_cols + (c)] \
                                 : (double)(m)->u.data_int[(size_t)(r) * (m)->num_cols + (c)])
    if (keys) { /* left keys, then right keys shifted by the left width */
        kd.num_rows = keys_left->num_rows + keys_right->num_rows;
        kd.num_cols = keys_left->num_cols;
        if (keys_right->num_cols != kd.num_cols || mat_resize(&kd)) goto done;
        for (i = 0; i < kd.num_rows; i++)
            for (j = 0; j < kd.num_cols; j++) {
                const int r = i < keys_left->num_rows;
                const Mat_rm *m = r ? keys_left : keys_right;
                const int row = r ? i : i - keys_left->num_rows;
                kd.u.data_double[(size_t)i * kd.num_cols + j] =
                    MAT_AT(m, row, j) + ((!r && j == 0) ? right_pad : 0.0);
            }
        if (draw_points(&kd, &carg->nx, 1, keys)) goto done;
    }
    if (lines) {
        md.num_rows = match_right->num_rows;
        md.num_cols = match_right->num_cols;
        if (mat_resize(&md)) goto done;
        for (i = 0; i < md.num_rows; i++)
            for (j = 0; j < md.num_cols; j++)
                md.u.data_double[(size_t)i * md.num_cols + j] =
                    MAT_AT(match_right, i, j) + (j == 0 ? right_pad : 0.0);
        if (draw_lines(match_left, &md, &carg->nx, lines)) goto done;
    }
#undef MAT_AT
    rc = SIFT3D_SUCCESS;
done:
    free(ctmp.data);
    free(lp.data);
    free(rp.data);
    free(kd.u.data_double);
    free(md.u.data_double);
    return rc;
}
