// Minimal C++ host program over include/pcr_b200.hpp (the mirror of the reference's Rust API).
//   g++ -std=c++17 -Iinclude examples/cpp_host_example.cpp -Lpointclouds_rs_b200/lib -lpcr_b200 \
//       -Wl,-rpath,$PWD/pointclouds_rs_b200/lib -o /tmp/cpp_host_example
#include <cstdio>
#include <random>

#include "pcr_b200.hpp"

int main() {
    try {
        std::mt19937 rng(42);
        std::uniform_real_distribution<float> u(0.f, 10.f);
        std::vector<float> x(5000), y(5000), z(5000);
        for (size_t i = 0; i < x.size(); i++) { x[i] = u(rng); y[i] = u(rng); z[i] = 0.01f * u(rng); }
        x.push_back(100.f); y.push_back(100.f); z.push_back(100.f);  // one far outlier
        pcr::PointCloud cloud = pcr::PointCloud::from_xyz(x, y, z);
        pcr::PointCloud clean = pcr::statistical_outlier_removal(cloud, 10, 1.0f);
        pcr::Normals n = pcr::estimate_normals(clean, 20);
        std::printf("kept %zu of %zu, first normal (%g, %g, %g)\n", clean.len(), cloud.len(), n.nx[0], n.ny[0], n.nz[0]);
        // the same pipeline without leaving the device: one upload, one download
        pcr::DeviceCloud dev = pcr::DeviceCloud::upload(cloud);
        pcr::DeviceCloud out = dev.voxel_downsample(0.05f).statistical_outlier_removal(10, 1.0f).estimate_normals(20);
        auto clusters = out.euclidean_cluster(0.5f, 10, 100000);
        pcr::PointCloud host = out.download();
        std::printf("device pipeline: %zu points with normals, %zu clusters (largest %zu)\n", host.len(), clusters.size(),
                    clusters.empty() ? 0 : clusters[0].size());
    } catch (const pcr::Error &e) {
        std::printf("pcr error %d: %s\n", e.code, e.what());  // e.g. 6 = no CUDA device (there is no CPU fallback)
        return e.code == PCR_ERR_NO_DEVICE ? 0 : 1;
    }
    return 0;
}
