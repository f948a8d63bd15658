/*
 * pcr_oracle.h -- CPU oracle for the pcrs KNN hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference algorithms (nahomes-15/pointclouds-rs):
 * every function cites the reference file:line it follows.  It exists to CHECK the CUDA
 * path (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference leg).
 * Nothing under pointclouds_rs_b200/ may call, link or import it.
 *
 * Parity status: the reference itself (Rust + kiddo 5.2.4 + nalgebra 0.33.2) cannot be
 * built in this image (no cargo/rustc, crates not vendored).  The oracle is pinned against
 * every known-answer test the reference holds for this path (tests/test_oracle_reference_kats.py)
 * and cross-checked against scipy.spatial.cKDTree; exact KNN index lists / SOR masks on
 * realistic inputs are NOT pinned by the reference's own tests ("parity unpinned" for those:
 * the oracle is the pin).  Where kiddo leaves behaviour unspecified (order among equal
 * squared distances, choice at the k-th boundary) the oracle defines it: ascending
 * (d^2, index).
 *
 * Arithmetic rules (build with -ffp-contract=off, SSE2 f32):
 *   d^2 = ((dx*dx) + (dy*dy)) + (dz*dz), each op individually rounded in f32
 *   (kiddo SquaredEuclidean over f32, call sites crates/spatial/src/kdtree.rs:70,93,121-123).
 */
#ifndef PCR_ORACLE_H
#define PCR_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_tree orc_tree;

/* crates/spatial/src/kdtree.rs:25-44 (KdTree::build).  Non-finite points are left out of the
 * index (engine rule; the reference never builds a tree containing NaN). */
orc_tree *orc_tree_build(const float *x, const float *y, const float *z, size_t n);
void orc_tree_free(orc_tree *t);
size_t orc_tree_len(const orc_tree *t); /* kdtree.rs:47-49: number of points of the cloud */

/* kdtree.rs:64-80 (knn) / :87-96 (knn_indices).  Returns the number of results (<= k).
 * idx/dist may be NULL.  dist = sqrt(d^2) (Euclidean), ascending, ties by index. */
size_t orc_tree_knn(const orc_tree *t, const float q[3], size_t k, uint32_t *idx, float *dist);
/* same search, brute force over all points: validates the kd-tree. */
size_t orc_knn_brute(const float *x, const float *y, const float *z, size_t n, const float q[3],
                     size_t k, uint32_t *idx, float *dist);

/* Batched form: row-major nq x k outputs, rows padded with idx=UINT32_MAX / dist=INFINITY,
 * counts[q] = number of valid entries.  threads<=1: serial. */
void orc_knn_batch(const orc_tree *t, const float *qx, const float *qy, const float *qz, size_t nq,
                   size_t k, uint32_t *idx, float *dist, uint32_t *counts, int threads);

/* kdtree.rs:105-135 (radius_search, index-sorted).  Returns count; writes at most cap indices. */
size_t orc_tree_radius_search(const orc_tree *t, const float q[3], float radius, uint32_t *idx,
                              size_t cap);
/* kdtree.rs:142-163 (radius_search_unsorted) .len() */
size_t orc_tree_radius_count(const orc_tree *t, const float q[3], float radius);
void orc_radius_count_batch(const orc_tree *t, const float *qx, const float *qy, const float *qz,
                            size_t nq, float radius, uint32_t *counts, int threads);

/* crates/filters/src/statistical_outlier.rs:4-69.  keep[i] in {0,1}; mean_d (n floats) and
 * stats[3] = {global_mean, global_stddev, threshold} may be NULL.  Returns kept count.
 * `threads` parallelises only the per-point KNN (the reference is serial, :19); the folds at
 * :53-60 are always sequential f32. */
size_t orc_sor(const float *x, const float *y, const float *z, size_t n, size_t k, float std_mul,
               uint8_t *keep, float *mean_d, float *stats, int threads);

/* crates/filters/src/radius_outlier.rs:4-18 */
size_t orc_ror(const float *x, const float *y, const float *z, size_t n, float radius,
               size_t min_neighbors, uint8_t *keep, int threads);

/* crates/normals/src/estimate.rs:19-124 (+ :139-238).  threads mirrors rayon par_iter (:42-44). */
void orc_normals(const float *x, const float *y, const float *z, size_t n, size_t k,
                 const float viewpoint[3], float *nx, float *ny, float *nz, int threads);
/* estimate.rs:139-238 exposed for unit tests */
void orc_smallest_eigenvector_3x3(float a00, float a01, float a02, float a11, float a12, float a22,
                                  float out[3]);

/* crates/registration/src/icp.rs:8-11, 112-118 */
typedef struct {
    float rotation[9]; /* row-major 3x3 */
    float translation[3];
} orc_transform;

typedef struct {
    orc_transform transform;
    float fitness;
    float rmse;
    int converged;
    size_t num_iterations;
} orc_icp_result;

typedef struct {
    size_t max_iterations;             /* icp.rs:95-109 defaults: 50, 1e-5, INF */
    float tolerance;
    float max_correspondence_distance;
} orc_icp_params;

/* icp.rs:77-92 / :39-47 */
void orc_apply_transform(const float *x, const float *y, const float *z, size_t n,
                         const orc_transform *t, float *ox, float *oy, float *oz);
/* icp.rs:52-73: apply `a` first, then `b` */
void orc_compose(const orc_transform *a, const orc_transform *b, orc_transform *out);

/* crates/registration/src/correspondence.rs:16-39.  Outputs sized ns; returns count. */
size_t orc_find_correspondences(const float *sx, const float *sy, const float *sz, size_t ns,
                                const orc_tree *target, float max_distance, uint32_t *src_idx,
                                uint32_t *tgt_idx, float *dist, int threads);

/* icp.rs:125-206 */
void orc_icp_point_to_point(const float *sx, const float *sy, const float *sz, size_t ns,
                            const float *tx, const float *ty, const float *tz, size_t nt,
                            const orc_icp_params *p, orc_icp_result *out, int threads);
/* icp_plane.rs:20-97; returns 0 ok, 1 = NormalsMismatch (:27-32) */
int orc_icp_point_to_plane(const float *sx, const float *sy, const float *sz, size_t ns,
                           const float *tx, const float *ty, const float *tz, size_t nt,
                           const float *nx, const float *ny, const float *nz, size_t nn,
                           const orc_icp_params *p, orc_icp_result *out, int threads);

/* crates/segmentation/src/euclidean_cluster.rs:96-187.  CSR output in the reference's order (size
 * descending, then smallest index ascending; indices ascending inside a cluster): offsets needs
 * n + 1 entries, indices n.  Returns the number of clusters.  Non-finite points are singleton
 * components (:112-119 leaves them out of the grid, :164-168 still lists them).  The dx*dx+dy*dy+dz*dz
 * of :147-150 is the same left-to-right f32 expression as kiddo's d^2. */
size_t orc_euclidean_cluster(const float *x, const float *y, const float *z, size_t n, float distance_threshold,
                             size_t min_size, size_t max_size, uint32_t *offsets, uint32_t *indices);
/* tests/cluster_differential.rs:13-82 (the reference's own brute-force checker), same output form */
size_t orc_euclidean_cluster_brute(const float *x, const float *y, const float *z, size_t n, float distance_threshold,
                                   size_t min_size, size_t max_size, uint32_t *offsets, uint32_t *indices);

/* crates/segmentation/src/ransac_plane.rs:56-129 for GIVEN samples: the sampling (:75-78, :140-163) uses
 * StdRng = ChaCha12, whose stream cannot be reproduced here, so `samples` holds the m index triples exactly as
 * sample_three_distinct would hand them over.  Both execution paths are followed: sequential with the adaptive
 * early exit (:95-121) and, for n >= 10000 and >= 16 samples, the order-preserving parallel reduction (:82-93).
 * model = {nx, ny, nz, d}; inliers sized n; returns the inlier count. */
size_t orc_ransac_plane_samples(const float *x, const float *y, const float *z, size_t n, float threshold,
                                const uint32_t *samples, size_t m, float model[4], uint32_t *inliers);

/* crates/filters/src/voxel_downsample.rs:12-65.  Outputs sized n; returns voxel count,
 * or (size_t)-1 if voxel_size is not finite / <= 0 (the reference panics, :13-16). */
size_t orc_voxel_downsample(const float *x, const float *y, const float *z, size_t n,
                            float voxel_size, float *ox, float *oy, float *oz);

/* crates/io/src/pcd.rs:202-234 (ASCII body only).  Returns point count or -1 on IO error. */
long orc_read_pcd_ascii(const char *path, float *x, float *y, float *z, size_t cap);

/* Uniform-grid ring-search MODEL (scalar C restatement of the CUDA engine's search rule, used to
 * validate the termination bound on CPU and to tune the cell size).  Same results as
 * orc_tree_knn by construction; stats[0]+=candidates examined, stats[1]+=rings, stats[2]+=runs (contiguous cell ranges
 * looked up), stats[3]+=32-wide batches. stats has 4 entries. */
void orc_grid_knn_model(const float *x, const float *y, const float *z, size_t n, const float *qx,
                        const float *qy, const float *qz, size_t nq, size_t k, float cell,
                        uint32_t *idx, float *dist, uint32_t *counts, double *stats);

#ifdef __cplusplus
}
#endif
#endif
