// cubagraph_reader.h -- reads a flat .cubagraph file (see cuda-bundle-adjustment_b200/graphio.py) into graph objects of the
// drop-in API, the way the reference's sample reads its JSON (samples/sample_ba_from_file.cpp:91-157 of the reference).
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <cuda_bundle_adjustment.h>

struct Storage {
	std::vector<std::unique_ptr<cuba::PoseVertex>> poses;
	std::vector<std::unique_ptr<cuba::LandmarkVertex>> landmarks;
	std::vector<std::unique_ptr<cuba::MonoEdge>> mono;
	std::vector<std::unique_ptr<cuba::StereoEdge>> stereo;
};

template <typename T>
static inline std::vector<T> readArray(FILE* f, size_t n)
{
	std::vector<T> v(n);
	if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
	return v;
}

static inline cuba::CudaBundleAdjustment::Ptr readGraph(const std::string& path, Storage& st)
{
	FILE* f = fopen(path.c_str(), "rb");
	if (!f) { fprintf(stderr, "cannot open %s\n", path.c_str()); exit(2); }
	char magic[8];
	if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "CUBAGRF1", 8) != 0) { fprintf(stderr, "bad magic\n"); exit(2); }
	const auto n = readArray<int64_t>(f, 4);
	const size_t nP = n[0], nL = n[1], nM = n[2], nS = n[3];
	const auto pid = readArray<int32_t>(f, nP), pfix = readArray<int32_t>(f, nP);
	const auto q = readArray<double>(f, 4 * nP), t = readArray<double>(f, 3 * nP), cam = readArray<double>(f, 5 * nP);
	const auto lid = readArray<int32_t>(f, nL), lfix = readArray<int32_t>(f, nL);
	const auto Xw = readArray<double>(f, 3 * nL);
	const auto mP = readArray<int32_t>(f, nM), mL = readArray<int32_t>(f, nM);
	const auto mMeas = readArray<double>(f, 2 * nM), mInfo = readArray<double>(f, nM);
	const auto sP = readArray<int32_t>(f, nS), sL = readArray<int32_t>(f, nS);
	const auto sMeas = readArray<double>(f, 3 * nS), sInfo = readArray<double>(f, nS);
	fclose(f);

	auto optimizer = cuba::CudaBundleAdjustment::create();
	for (size_t i = 0; i < nP; i++) {
		cuba::CameraParams c;
		c.fx = cam[5 * i]; c.fy = cam[5 * i + 1]; c.cx = cam[5 * i + 2]; c.cy = cam[5 * i + 3]; c.bf = cam[5 * i + 4];
		cuba::PoseVertex::Quaternion qq;
		for (int k = 0; k < 4; k++) qq.coeffs().data()[k] = q[4 * i + k];
		cuba::PoseVertex::Translation tt;
		for (int k = 0; k < 3; k++) tt.data()[k] = t[3 * i + k];
		st.poses.emplace_back(new cuba::PoseVertex(pid[i], qq, tt, c, pfix[i] != 0));
		optimizer->addPoseVertex(st.poses.back().get());
	}
	for (size_t i = 0; i < nL; i++) {
		cuba::LandmarkVertex::Point3D X;
		for (int k = 0; k < 3; k++) X.data()[k] = Xw[3 * i + k];
		st.landmarks.emplace_back(new cuba::LandmarkVertex(lid[i], X, lfix[i] != 0));
		optimizer->addLandmarkVertex(st.landmarks.back().get());
	}
	for (size_t i = 0; i < nM; i++) {
		cuba::MonoEdge::Measurement m;
		m.data()[0] = mMeas[2 * i]; m.data()[1] = mMeas[2 * i + 1];
		st.mono.emplace_back(new cuba::MonoEdge(m, mInfo[i], optimizer->poseVertex(mP[i]), optimizer->landmarkVertex(mL[i])));
		optimizer->addMonocularEdge(st.mono.back().get());
	}
	for (size_t i = 0; i < nS; i++) {
		cuba::StereoEdge::Measurement m;
		for (int k = 0; k < 3; k++) m.data()[k] = sMeas[3 * i + k];
		st.stereo.emplace_back(new cuba::StereoEdge(m, sInfo[i], optimizer->poseVertex(sP[i]), optimizer->landmarkVertex(sL[i])));
		optimizer->addStereoEdge(st.stereo.back().get());
	}
	return optimizer;
}

