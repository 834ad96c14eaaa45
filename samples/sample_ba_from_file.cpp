// sample_ba_from_file.cpp -- the reference's benchmark protocol on the drop-in API.
//
// Mirrors samples/sample_ba_from_file.cpp of the reference (graph -> warm-up initialize()+optimize(1)
// -> timed initialize()+optimize(10) -> time profile + chi2 per iteration, :34-89,159-161) but reads the
// flat .cubagraph format (see cuda-bundle-adjustment_b200/graphio.py) instead of OpenCV's JSON reader,
// and can print one JSON object for the tests (--json).  Built by tests/test_cpp_api.py and by hand:
//   g++ -std=c++17 -O2 -Iinclude samples/sample_ba_from_file.cpp -Lcuda-bundle-adjustment_b200 -lcuba_b200 \
//       -Wl,-rpath,$PWD/cuda-bundle-adjustment_b200 -o sample_ba_from_file
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
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
static std::vector<T> readArray(FILE* f, size_t n)
{
	std::vector<T> v(n);
	if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
	return v;
}

static cuba::CudaBundleAdjustment::Ptr readGraph(const std::string& path, Storage& st)
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

int main(int argc, char** argv)
{
	if (argc < 2) { printf("Usage: sample_ba_from_file input.cubagraph [--json] [--huber] [--iters N] [--no-warmup]\n"); return 0; }
	bool json = false, huber = false, warmup = true;
	int iters = 10;
	for (int i = 2; i < argc; i++) {
		if (!strcmp(argv[i], "--json")) json = true;
		else if (!strcmp(argv[i], "--huber")) huber = true;
		else if (!strcmp(argv[i], "--no-warmup")) warmup = false;
		else if (!strcmp(argv[i], "--iters") && i + 1 < argc) iters = atoi(argv[++i]);
	}
	Storage st;
	auto optimizer = readGraph(argv[1], st);
	if (huber) {  // the g2o comparison sample's kernels (reference samples/sample_comparison_with_g2o.cpp:195-200)
		optimizer->setRobustKernels(cuba::RobustKernelType::HUBER, std::sqrt(5.991), cuba::EdgeType::MONOCULAR);
		optimizer->setRobustKernels(cuba::RobustKernelType::HUBER, std::sqrt(7.815), cuba::EdgeType::STEREO);
	}
	if (warmup) { optimizer->initialize(); optimizer->optimize(1); }   // writes its result back, like the reference

	const auto t0 = std::chrono::steady_clock::now();
	optimizer->initialize();
	optimizer->optimize(iters);
	const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

	if (json) {
		printf("{\"nposes\": %zu, \"nlandmarks\": %zu, \"nedges\": %zu, \"seconds\": %.6f, \"chi2\": [",
			optimizer->nposes(), optimizer->nlandmarks(), optimizer->nedges(), sec);
		const auto& s = optimizer->batchStatistics();
		for (size_t i = 0; i < s.size(); i++) printf("%s%.17g", i ? ", " : "", s[i].chi2);
		double chiSum = 0;
		for (auto& e : st.mono) chiSum += optimizer->chiSquared(e.get());
		for (auto& e : st.stereo) chiSum += optimizer->chiSquared(e.get());
		printf("], \"sum_edge_chi2\": %.17g, \"t_last\": [%.17g, %.17g, %.17g], \"profile\": {", chiSum,
			st.poses.back()->t.data()[0], st.poses.back()->t.data()[1], st.poses.back()->t.data()[2]);
		bool first = true;
		for (const auto& kv : optimizer->timeProfile()) { printf("%s\"%s\": %.6f", first ? "" : ", ", kv.first.c_str(), kv.second); first = false; }
		printf("}}\n");
		return 0;
	}
	printf("=== Graph size : \nnum poses      : %zu\nnum landmarks  : %zu\nnum edges      : %zu\n\n", optimizer->nposes(), optimizer->nlandmarks(), optimizer->nedges());
	printf("=== Processing time : \nBA total : %.4f[sec]\n\n", sec);
	for (const auto& kv : optimizer->timeProfile()) printf("%-30s : %8.2f[msec]\n", kv.first.c_str(), 1e3 * kv.second);
	printf("\n=== Objective function value : \n");
	for (const auto& s : optimizer->batchStatistics()) printf("iter: %2d, chi2: %.1f\n", s.iteration + 1, s.chi2);
	return 0;
}
