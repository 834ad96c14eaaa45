// dynamic_graph_driver.cpp -- TEST DRIVER for the dynamic graph semantics of the drop-in API (reference
// src/cuda_bundle_adjustment.cpp:677-781: add*/remove*/re-initialize()/repeated optimize()/chiSquared()), which the reference
// itself never tests.  Executes a ';'-separated op list on a .cubagraph and prints one JSON object; tests/test_dynamic_graph.py
// mirrors the same ops on the graph arrays and checks every optimize() against the CPU oracle.
//   init | opt:N | rmpose:ID | rmlm:ID | rmedge:m:K | rmedge:s:K | addedge:m:K | addedge:s:K | outliers:T | fixp:ID | fixl:ID
#include <cmath>
#include <sstream>

#include "../../samples/cubagraph_reader.h"

int main(int argc, char** argv)
{
	if (argc < 4) { fprintf(stderr, "usage: dynamic_graph_driver graph.cubagraph ops dump.bin [--huber]\n"); return 2; }
	Storage st;
	auto opt = readGraph(argv[1], st);
	if (argc > 4 && !strcmp(argv[4], "--huber")) {
		opt->setRobustKernels(cuba::RobustKernelType::HUBER, std::sqrt(5.991), cuba::EdgeType::MONOCULAR);
		opt->setRobustKernels(cuba::RobustKernelType::HUBER, std::sqrt(7.815), cuba::EdgeType::STEREO);
	}
	std::stringstream ss(argv[2]);
	std::string op;
	printf("{\"steps\": [");
	bool first = true;
	while (std::getline(ss, op, ';')) {
		if (op.empty()) continue;
		std::vector<std::string> f;
		{ std::stringstream s2(op); std::string x; while (std::getline(s2, x, ':')) f.push_back(x); }
		printf("%s{\"op\": \"%s\"", first ? "" : ", ", op.c_str());
		first = false;
		if (f[0] == "init") opt->initialize();
		else if (f[0] == "opt") {
			const size_t before = opt->batchStatistics().size();
			opt->optimize(atoi(f[1].c_str()));
			const auto& s = opt->batchStatistics();
			printf(", \"stats_before\": %zu, \"chi2\": [", before);
			for (size_t i = 0; i < s.size(); i++) printf("%s%.17g", i ? ", " : "", s[i].chi2);
			double sum = 0;
			for (auto& e : st.mono) sum += opt->chiSquared(e.get());
			for (auto& e : st.stereo) sum += opt->chiSquared(e.get());
			printf("], \"sum_edge_chi2\": %.17g", sum);
		}
		else if (f[0] == "rmpose") opt->removePoseVertex(opt->poseVertex(atoi(f[1].c_str())));
		else if (f[0] == "rmlm") opt->removeLandmarkVertex(opt->landmarkVertex(atoi(f[1].c_str())));
		else if (f[0] == "rmedge") { const size_t k = atol(f[2].c_str()); if (f[1] == "m") opt->removeEdge(st.mono[k].get()); else opt->removeEdge(st.stereo[k].get()); }
		else if (f[0] == "addedge") { const size_t k = atol(f[2].c_str()); if (f[1] == "m") opt->addMonocularEdge(st.mono[k].get()); else opt->addStereoEdge(st.stereo[k].get()); }
		else if (f[0] == "fixp") opt->poseVertex(atoi(f[1].c_str()))->fixed = true;
		else if (f[0] == "fixl") opt->landmarkVertex(atoi(f[1].c_str()))->fixed = true;
		else if (f[0] == "outliers") {
			// the ORB-SLAM pattern: drop every edge whose chi2 exceeds a threshold, then optimise again
			const double T = atof(f[1].c_str());
			size_t n = 0;
			for (auto& e : st.mono) if (opt->chiSquared(e.get()) > T) { opt->removeEdge(e.get()); n++; }
			for (auto& e : st.stereo) if (opt->chiSquared(e.get()) > T) { opt->removeEdge(e.get()); n++; }
			printf(", \"removed\": %zu", n);
		}
		else { fprintf(stderr, "unknown op %s\n", op.c_str()); return 2; }
		printf(", \"nposes\": %zu, \"nlandmarks\": %zu, \"nedges\": %zu}", opt->nposes(), opt->nlandmarks(), opt->nedges());
	}
	printf("]}\n");
	bool threw = false;
	try { opt->poseVertex(-12345); } catch (const std::out_of_range&) { threw = true; }
	if (!threw) { fprintf(stderr, "poseVertex(unknown id) did not throw std::out_of_range\n"); return 3; }
	FILE* fo = fopen(argv[3], "wb");
	if (!fo) return 2;
	for (auto& v : st.poses) fwrite(v->q.coeffs().data(), sizeof(double), 4, fo);
	for (auto& v : st.poses) fwrite(v->t.data(), sizeof(double), 3, fo);
	for (auto& v : st.landmarks) fwrite(v->Xw.data(), sizeof(double), 3, fo);
	fclose(fo);
	return 0;
}
