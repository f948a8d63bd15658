// pcr_b200.hpp -- header-only C++17 host layer over the C ABI (pcr_b200.h) that mirrors the Rust API
// of pcrs for the KNN path: same names, argument meaning and degenerate-input behaviour, so that the
// reference's own tests read the same against it.  (The reference host language is Rust; there is no
// Rust toolchain in this image, see DESIGN.md section 1 and INTEGRATION.md for the Rust binding.)
//
//   pcr::PointCloud / Normals            crates/core/src/cloud.rs:4-18, :103-162
//   pcr::KdTree                          crates/spatial/src/kdtree.rs:14-164 (batched queries)
//   pcr::statistical_outlier_removal     crates/filters/src/statistical_outlier.rs:4
//   pcr::radius_outlier_removal          crates/filters/src/radius_outlier.rs:4
//   pcr::estimate_normals[_with_viewpoint]  crates/normals/src/estimate.rs:13,19
//   pcr::find_correspondences            crates/registration/src/correspondence.rs:16
//   pcr::apply_transform, RigidTransform crates/registration/src/icp.rs:8-92
//   pcr::icp_point_to_point / _plane     crates/registration/src/icp.rs:125, icp_plane.rs:20
//   pcr::voxel_downsample                crates/filters/src/voxel_downsample.rs:12
//   pcr::euclidean_cluster               crates/segmentation/src/euclidean_cluster.rs:96
//   pcr::ransac_plane_samples            crates/segmentation/src/ransac_plane.rs:56 (after its sampling step)
//   pcr::DeviceCloud                     the same calls on a cloud that stays in HBM between steps
#pragma once

#include <cmath>
#include <cstdint>
#include <limits>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "pcr_b200.h"

namespace pcr {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

struct Normals {
    std::vector<float> nx, ny, nz;
};

struct PointCloud {
    std::vector<float> x, y, z;
    std::optional<Normals> normals;

    static PointCloud from_xyz(std::vector<float> x, std::vector<float> y, std::vector<float> z) {
        if (x.size() != y.size() || x.size() != z.size()) throw std::invalid_argument("x, y and z must have same length");
        PointCloud c;
        c.x = std::move(x);
        c.y = std::move(y);
        c.z = std::move(z);
        return c;
    }
    size_t len() const { return x.size(); }
    bool is_empty() const { return x.empty(); }
    // cloud.rs:103-140: gathers the kept indices, preserving order (and normals if present)
    PointCloud select(const std::vector<size_t> &indices) const {
        PointCloud o;
        o.x.reserve(indices.size());
        o.y.reserve(indices.size());
        o.z.reserve(indices.size());
        for (size_t i : indices) {
            if (i >= len()) throw std::out_of_range("index out of bounds in select");
            o.x.push_back(x[i]);
            o.y.push_back(y[i]);
            o.z.push_back(z[i]);
        }
        if (normals) {
            Normals n;
            for (size_t i : indices) {
                n.nx.push_back(normals->nx[i]);
                n.ny.push_back(normals->ny[i]);
                n.nz.push_back(normals->nz[i]);
            }
            o.normals = std::move(n);
        }
        return o;
    }
};

class Context {
  public:
    explicit Context(int device = 0) {
        if (int s = pcr_ctx_create(device, &ctx_)) throw Error(s, pcr_last_error(nullptr));
    }
    ~Context() { pcr_ctx_destroy(ctx_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    pcr_ctx *get() const { return ctx_; }
    void check(int status) const {
        if (status != PCR_OK) throw Error(status, pcr_last_error(ctx_));
    }

  private:
    pcr_ctx *ctx_ = nullptr;
};

inline Context &default_context() {
    static thread_local Context ctx(0);
    return ctx;
}

class KdTree {
  public:
    static KdTree build(const PointCloud &cloud, size_t k_hint = 0, Context &ctx = default_context()) {
        KdTree t;
        t.ctx_ = &ctx;
        ctx.check(pcr_index_build(ctx.get(), cloud.x.data(), cloud.y.data(), cloud.z.data(), cloud.len(), k_hint, &t.ix_));
        return t;
    }
    KdTree(KdTree &&o) noexcept : ctx_(o.ctx_), ix_(o.ix_) { o.ix_ = nullptr; }
    KdTree(const KdTree &) = delete;
    ~KdTree() { pcr_index_free(ix_); }
    size_t len() const { return pcr_index_len(ix_); }
    bool is_empty() const { return len() == 0; }
    pcr_index *get() const { return ix_; }

    // knn for a batch: row q of idx/dist holds counts[q] valid entries, ascending (distance, index)
    void knn(const std::vector<float> &qx, const std::vector<float> &qy, const std::vector<float> &qz, size_t k,
             std::vector<uint32_t> &idx, std::vector<float> &dist, std::vector<uint32_t> &counts) const {
        const size_t nq = qx.size();
        idx.assign(nq * k, UINT32_MAX);
        dist.assign(nq * k, std::numeric_limits<float>::infinity());
        counts.assign(nq, 0);
        ctx_->check(pcr_knn(ix_, qx.data(), qy.data(), qz.data(), nq, k, idx.data(), dist.data(), counts.data()));
    }
    // single-query conveniences with the reference's exact signatures (kdtree.rs:64, :105)
    std::pair<std::vector<size_t>, std::vector<float>> knn(const float q[3], size_t k) const {
        std::vector<uint32_t> idx(k ? k : 1), cnt(1);
        std::vector<float> dist(k ? k : 1);
        ctx_->check(pcr_knn(ix_, &q[0], &q[1], &q[2], 1, k, idx.data(), dist.data(), cnt.data()));
        return {std::vector<size_t>(idx.begin(), idx.begin() + cnt[0]), std::vector<float>(dist.begin(), dist.begin() + cnt[0])};
    }
    std::vector<size_t> radius_search(const float q[3], float radius) const {
        uint64_t off[2] = {0, 0};
        size_t total = 0;
        std::vector<uint32_t> idx(64);
        int s = pcr_radius_search(ix_, &q[0], &q[1], &q[2], 1, radius, off, idx.data(), idx.size(), &total);
        if (s == PCR_ERR_CAPACITY) {
            idx.resize(total);
            s = pcr_radius_search(ix_, &q[0], &q[1], &q[2], 1, radius, off, idx.data(), idx.size(), &total);
        }
        ctx_->check(s);
        return std::vector<size_t>(idx.begin(), idx.begin() + total);
    }

  private:
    KdTree() = default;
    Context *ctx_ = nullptr;
    pcr_index *ix_ = nullptr;
};

inline PointCloud statistical_outlier_removal(const PointCloud &cloud, size_t k, float std_mul, Context &ctx = default_context()) {
    std::vector<uint8_t> keep(cloud.len() ? cloud.len() : 1);
    size_t kept = 0;
    ctx.check(pcr_sor(ctx.get(), cloud.x.data(), cloud.y.data(), cloud.z.data(), cloud.len(), k, std_mul, keep.data(), &kept, nullptr,
                      nullptr));
    std::vector<size_t> idx;
    idx.reserve(kept);
    for (size_t i = 0; i < cloud.len(); i++)
        if (keep[i]) idx.push_back(i);
    return cloud.select(idx);  // statistical_outlier.rs:68
}

inline PointCloud radius_outlier_removal(const PointCloud &cloud, float radius, size_t min_neighbors, Context &ctx = default_context()) {
    std::vector<uint8_t> keep(cloud.len() ? cloud.len() : 1);
    size_t kept = 0;
    ctx.check(pcr_radius_outlier(ctx.get(), cloud.x.data(), cloud.y.data(), cloud.z.data(), cloud.len(), radius, min_neighbors,
                                 keep.data(), &kept));
    std::vector<size_t> idx;
    for (size_t i = 0; i < cloud.len(); i++)
        if (keep[i]) idx.push_back(i);
    return cloud.select(idx);
}

inline Normals estimate_normals_with_viewpoint(const PointCloud &cloud, size_t k, const float viewpoint[3],
                                               Context &ctx = default_context()) {
    Normals n;
    if (cloud.is_empty() || k == 0) return n;  // estimate.rs:25-31
    n.nx.resize(cloud.len());
    n.ny.resize(cloud.len());
    n.nz.resize(cloud.len());
    ctx.check(pcr_estimate_normals(ctx.get(), cloud.x.data(), cloud.y.data(), cloud.z.data(), cloud.len(), k, viewpoint, n.nx.data(),
                                   n.ny.data(), n.nz.data()));
    return n;
}
inline Normals estimate_normals(const PointCloud &cloud, size_t k, Context &ctx = default_context()) {
    const float origin[3] = {0.f, 0.f, 0.f};
    return estimate_normals_with_viewpoint(cloud, k, origin, ctx);
}

struct RigidTransform {
    float rotation[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    float translation[3] = {0, 0, 0};
    static RigidTransform identity() { return RigidTransform(); }
};

struct IcpParams {
    size_t max_iterations = 50;
    float tolerance = 1e-5f;
    float max_correspondence_distance = std::numeric_limits<float>::infinity();
};

struct IcpResult {
    RigidTransform transform;
    float fitness = 0.f, rmse = 0.f;
    bool converged = false;
    size_t num_iterations = 0;
};

struct Correspondence {
    size_t source_index, target_index;
    float distance;
};

inline std::vector<Correspondence> find_correspondences(const PointCloud &source, const KdTree &target_tree, float max_distance,
                                                        Context &ctx = default_context()) {
    const size_t ns = source.len();
    std::vector<uint32_t> si(ns ? ns : 1), ti(ns ? ns : 1);
    std::vector<float> dd(ns ? ns : 1);
    size_t m = 0;
    ctx.check(pcr_find_correspondences(target_tree.get(), source.x.data(), source.y.data(), source.z.data(), ns, max_distance, si.data(),
                                       ti.data(), dd.data(), &m));
    std::vector<Correspondence> out(m);
    for (size_t i = 0; i < m; i++) out[i] = {si[i], ti[i], dd[i]};
    return out;
}

inline PointCloud apply_transform(const PointCloud &cloud, const RigidTransform &t, Context &ctx = default_context()) {
    PointCloud o;  // xyz only, like icp.rs:91
    o.x.resize(cloud.len());
    o.y.resize(cloud.len());
    o.z.resize(cloud.len());
    ctx.check(pcr_apply_transform(ctx.get(), cloud.x.data(), cloud.y.data(), cloud.z.data(), cloud.len(), &t.rotation[0][0],
                                  t.translation, o.x.data(), o.y.data(), o.z.data()));
    return o;
}

namespace detail {
inline IcpResult to_result(const pcr_icp_result &r) {
    IcpResult o;
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) o.transform.rotation[i][j] = r.rotation[i * 3 + j];
        o.transform.translation[i] = r.translation[i];
    }
    o.fitness = r.fitness;
    o.rmse = r.rmse;
    o.converged = r.converged != 0;
    o.num_iterations = (size_t)r.num_iterations;
    return o;
}
}  // namespace detail

inline IcpResult icp_point_to_point(const PointCloud &source, const PointCloud &target, const IcpParams &p,
                                    Context &ctx = default_context()) {
    pcr_icp_params cp{p.max_iterations, p.tolerance, p.max_correspondence_distance};
    pcr_icp_result r;
    ctx.check(pcr_icp_point_to_point(ctx.get(), source.x.data(), source.y.data(), source.z.data(), source.len(), target.x.data(),
                                     target.y.data(), target.z.data(), target.len(), &cp, &r));
    return detail::to_result(r);
}

// Throws pcr::Error with code PCR_ERR_NORMALS_MISMATCH where the reference returns
// Err(IcpPlaneError::NormalsMismatch) (icp_plane.rs:27-32).
inline IcpResult icp_point_to_plane(const PointCloud &source, const PointCloud &target, const Normals &target_normals,
                                    const IcpParams &p, Context &ctx = default_context()) {
    pcr_icp_params cp{p.max_iterations, p.tolerance, p.max_correspondence_distance};
    pcr_icp_result r;
    ctx.check(pcr_icp_point_to_plane(ctx.get(), source.x.data(), source.y.data(), source.z.data(), source.len(), target.x.data(),
                                     target.y.data(), target.z.data(), target.len(), target_normals.nx.data(), target_normals.ny.data(),
                                     target_normals.nz.data(), target_normals.nx.size(), &cp, &r));
    return detail::to_result(r);
}

// crates/filters/src/voxel_downsample.rs:12-65 (the reference asserts on a bad voxel size: std::invalid_argument here)
inline PointCloud voxel_downsample(const PointCloud &cloud, float voxel_size, Context &ctx = default_context()) {
    if (!std::isfinite(voxel_size) || !(voxel_size > 0.f)) throw std::invalid_argument("voxel_size must be > 0 and finite");
    const size_t n = cloud.len();
    std::vector<float> ox(n ? n : 1), oy(n ? n : 1), oz(n ? n : 1);
    size_t m = 0;
    ctx.check(pcr_voxel_downsample(ctx.get(), cloud.x.data(), cloud.y.data(), cloud.z.data(), n, voxel_size, ox.data(), oy.data(), oz.data(), &m));
    ox.resize(m);
    oy.resize(m);
    oz.resize(m);
    return PointCloud::from_xyz(std::move(ox), std::move(oy), std::move(oz));
}

// crates/segmentation/src/euclidean_cluster.rs:96-187
inline std::vector<std::vector<size_t>> euclidean_cluster(const PointCloud &cloud, float distance_threshold, size_t min_size,
                                                          size_t max_size, Context &ctx = default_context()) {
    const size_t n = cloud.len();
    std::vector<uint32_t> off(n + 1), idx(n ? n : 1);
    size_t nc = 0;
    ctx.check(pcr_euclidean_cluster(ctx.get(), cloud.x.data(), cloud.y.data(), cloud.z.data(), n, distance_threshold, min_size, max_size,
                                    off.data(), idx.data(), &nc));
    std::vector<std::vector<size_t>> out(nc);
    for (size_t c = 0; c < nc; c++) out[c].assign(idx.begin() + off[c], idx.begin() + off[c + 1]);
    return out;
}

// crates/segmentation/src/ransac_plane.rs:7-31
struct PlaneModel {
    float normal[3] = {0.f, 0.f, 1.f};
    float d = 0.f;
};

// ransac_plane_seeded (ransac_plane.rs:56-129) for index triples drawn by the caller (the reference draws them
// from StdRng, :75-78; a Rust host keeps doing that and passes them on)
inline std::pair<PlaneModel, std::vector<size_t>> ransac_plane_samples(const PointCloud &cloud, float distance_threshold,
                                                                       const std::vector<uint32_t> &samples /* 3 per hypothesis */,
                                                                       Context &ctx = default_context()) {
    const size_t n = cloud.len();
    float model[4];
    std::vector<uint32_t> inl(n ? n : 1);
    size_t k = 0;
    ctx.check(pcr_ransac_plane_samples(ctx.get(), cloud.x.data(), cloud.y.data(), cloud.z.data(), n, distance_threshold, samples.data(),
                                       samples.size() / 3, model, inl.data(), &k));
    PlaneModel m;
    m.normal[0] = model[0];
    m.normal[1] = model[1];
    m.normal[2] = model[2];
    m.d = model[3];
    return {m, std::vector<size_t>(inl.begin(), inl.begin() + k)};
}

// A PointCloud that stays in HBM between the steps of a pipeline (one upload, one download).
class DeviceCloud {
  public:
    static DeviceCloud upload(const PointCloud &cloud, Context &ctx = default_context()) {
        pcr_cloud *h = nullptr;
        ctx.check(pcr_cloud_upload(ctx.get(), cloud.x.data(), cloud.y.data(), cloud.z.data(), cloud.len(), &h));
        return DeviceCloud(h, &ctx);
    }
    // x | y | z rows of one host block, `stride` floats apart: one strided transfer
    static DeviceCloud upload_block(const float *xyz, size_t stride, size_t n, Context &ctx = default_context(), bool wait = true) {
        pcr_cloud *h = nullptr;
        ctx.check(wait ? pcr_cloud_upload_block(ctx.get(), xyz, stride, n, &h) : pcr_cloud_upload_block_nowait(ctx.get(), xyz, stride, n, &h));
        return DeviceCloud(h, &ctx);
    }
    // x | y | z [| nx | ny | nz] rows into one host block
    void download_block(float *dst, size_t stride, bool with_normals) const {
        ctx_->check(pcr_cloud_download_block(h_, dst, stride, with_normals ? 1 : 0));
    }
    ~DeviceCloud() { pcr_cloud_free(h_); }
    DeviceCloud(DeviceCloud &&o) noexcept : h_(o.h_), ctx_(o.ctx_) { o.h_ = nullptr; }
    DeviceCloud &operator=(DeviceCloud &&o) noexcept {
        if (this != &o) {
            pcr_cloud_free(h_);
            h_ = o.h_;
            ctx_ = o.ctx_;
            o.h_ = nullptr;
        }
        return *this;
    }
    DeviceCloud(const DeviceCloud &) = delete;
    DeviceCloud &operator=(const DeviceCloud &) = delete;

    size_t len() const { return pcr_cloud_len(h_); }
    bool is_empty() const { return len() == 0; }
    bool has_normals() const { return pcr_cloud_has_normals(h_) != 0; }
    PointCloud download() const {
        const size_t n = len();
        PointCloud c;
        c.x.resize(n);
        c.y.resize(n);
        c.z.resize(n);
        ctx_->check(pcr_cloud_download(h_, c.x.data(), c.y.data(), c.z.data()));
        if (has_normals()) {
            Normals nr;
            nr.nx.resize(n);
            nr.ny.resize(n);
            nr.nz.resize(n);
            ctx_->check(pcr_cloud_download_normals(h_, nr.nx.data(), nr.ny.data(), nr.nz.data()));
            c.normals = std::move(nr);
        }
        return c;
    }
    DeviceCloud select(const std::vector<uint32_t> &indices) const { return make(&pcr_cloud_select, indices.data(), indices.size()); }
    DeviceCloud voxel_downsample(float voxel_size) const { return make(&pcr_cloud_voxel_downsample, voxel_size); }
    DeviceCloud statistical_outlier_removal(size_t k, float std_mul) const { return make(&pcr_cloud_statistical_outlier_removal, k, std_mul); }
    DeviceCloud radius_outlier_removal(float radius, size_t min_neighbors) const {
        return make(&pcr_cloud_radius_outlier_removal, radius, min_neighbors);
    }
    DeviceCloud estimate_normals(size_t k) const { return make(&pcr_cloud_estimate_normals, k, (const float *)nullptr); }
    DeviceCloud apply_transform(const RigidTransform &t) const {
        return make(&pcr_cloud_apply_transform, &t.rotation[0][0], (const float *)t.translation);
    }
    std::vector<std::vector<size_t>> euclidean_cluster(float distance_threshold, size_t min_size, size_t max_size) const {
        const size_t n = len();
        std::vector<uint32_t> off(n + 1), idx(n ? n : 1);
        size_t nc = 0;
        ctx_->check(pcr_cloud_euclidean_cluster(h_, distance_threshold, min_size, max_size, off.data(), idx.data(), &nc));
        std::vector<std::vector<size_t>> out(nc);
        for (size_t c = 0; c < nc; c++) out[c].assign(idx.begin() + off[c], idx.begin() + off[c + 1]);
        return out;
    }
    IcpResult icp_point_to_point(const DeviceCloud &target, const IcpParams &p) const {
        pcr_icp_params cp{p.max_iterations, p.tolerance, p.max_correspondence_distance};
        pcr_icp_result r;
        ctx_->check(pcr_cloud_icp_point_to_point(h_, target.h_, &cp, &r));
        return detail::to_result(r);
    }
    IcpResult icp_point_to_plane(const DeviceCloud &target, const IcpParams &p) const {
        pcr_icp_params cp{p.max_iterations, p.tolerance, p.max_correspondence_distance};
        pcr_icp_result r;
        ctx_->check(pcr_cloud_icp_point_to_plane(h_, target.h_, &cp, &r));
        return detail::to_result(r);
    }

  private:
    DeviceCloud(pcr_cloud *h, Context *c) : h_(h), ctx_(c) {}
    template <class Fn, class... A>
    DeviceCloud make(Fn fn, A... args) const {
        pcr_cloud *o = nullptr;
        ctx_->check(fn(h_, args..., &o));
        return DeviceCloud(o, ctx_);
    }
    pcr_cloud *h_ = nullptr;
    Context *ctx_ = nullptr;
};

}  // namespace pcr
