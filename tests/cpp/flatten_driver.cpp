// flatten_driver.cpp -- TEST DRIVER (needs no GPU) for the graph container of the drop-in class: executes a ';'-separated op list on
// a .cubagraph and, after every `init`, appends the flat arrays initialize() built (cuba_debug_dropin_problem) to a binary file:
//   int64[6] Pall numP Lall numL E2 E3, then q t cam Xw (double), idx2 (int32), meas2 omega2 (double), idx3 (int32), meas3 omega3 (double).
// tests/test_host_initialize.py mirrors the ops on the graph arrays and compares with graphio.flatten (reference
// src/cuda_bundle_adjustment.cpp:142-243).
//   init | rmpose:ID | rmlm:ID | rmedge:m|s:K | addedge:m|s:K | fixp:ID | fixl:ID | unfixl:ID | meas:m|s:K:DELTA | movel:ID:DELTA
#include <sstream>

#include <cuba_b200.h>

#include "../../samples/cubagraph_reader.h"

int main(int argc, char** argv)
{
	if (argc < 4) { fprintf(stderr, "usage: flatten_driver graph.cubagraph ops dump.bin\n"); return 2; }
	Storage st;
	auto opt = readGraph(argv[1], st);
	FILE* fo = fopen(argv[3], "wb");
	if (!fo) return 2;
	std::stringstream ss(argv[2]);
	std::string op;
	int ninit = 0;
	while (std::getline(ss, op, ';')) {
		if (op.empty()) continue;
		std::vector<std::string> f;
		{ std::stringstream s2(op); std::string x; while (std::getline(s2, x, ':')) f.push_back(x); }
		if (f[0] == "init") {
			opt->initialize();
			cuba_problem p;
			if (cuba_debug_dropin_problem(opt.get(), &p) != CUBA_OK) { fprintf(stderr, "no flat problem\n"); return 3; }
			const int64_t n[6] = { p.Pall, p.numP, p.Lall, p.numL, p.E2, p.E3 };
			fwrite(n, sizeof(int64_t), 6, fo);
			fwrite(p.q, sizeof(double), 4 * (size_t)p.Pall, fo); fwrite(p.t, sizeof(double), 3 * (size_t)p.Pall, fo);
			fwrite(p.cam, sizeof(double), 5 * (size_t)p.Pall, fo); fwrite(p.Xw, sizeof(double), 3 * (size_t)p.Lall, fo);
			fwrite(p.idx2, sizeof(int32_t), 2 * (size_t)p.E2, fo); fwrite(p.meas2, sizeof(double), 2 * (size_t)p.E2, fo); fwrite(p.omega2, sizeof(double), (size_t)p.E2, fo);
			fwrite(p.idx3, sizeof(int32_t), 2 * (size_t)p.E3, fo); fwrite(p.meas3, sizeof(double), 3 * (size_t)p.E3, fo); fwrite(p.omega3, sizeof(double), (size_t)p.E3, fo);
			ninit++;
		}
		else if (f[0] == "rmpose") opt->removePoseVertex(opt->poseVertex(atoi(f[1].c_str())));
		else if (f[0] == "rmlm") opt->removeLandmarkVertex(opt->landmarkVertex(atoi(f[1].c_str())));
		else if (f[0] == "rmedge") { const size_t k = atol(f[2].c_str()); if (f[1] == "m") opt->removeEdge(st.mono[k].get()); else opt->removeEdge(st.stereo[k].get()); }
		else if (f[0] == "addedge") { const size_t k = atol(f[2].c_str()); if (f[1] == "m") opt->addMonocularEdge(st.mono[k].get()); else opt->addStereoEdge(st.stereo[k].get()); }
		else if (f[0] == "fixp") opt->poseVertex(atoi(f[1].c_str()))->fixed = true;
		else if (f[0] == "fixl") opt->landmarkVertex(atoi(f[1].c_str()))->fixed = true;
		else if (f[0] == "unfixl") opt->landmarkVertex(atoi(f[1].c_str()))->fixed = false;
		else if (f[0] == "meas") {      // the caller edits a measurement in place between two initialize() calls
			const size_t k = atol(f[2].c_str()); const double d = atof(f[3].c_str());
			if (f[1] == "m") st.mono[k]->measurement.data()[0] += d; else st.stereo[k]->measurement.data()[2] += d;
		}
		else if (f[0] == "movel") opt->landmarkVertex(atoi(f[1].c_str()))->Xw.data()[1] += atof(f[2].c_str());
		else { fprintf(stderr, "unknown op %s\n", op.c_str()); return 2; }
	}
	fclose(fo);
	printf("{\"inits\": %d, \"nposes\": %zu, \"nlandmarks\": %zu, \"nedges\": %zu}\n", ninit, opt->nposes(), opt->nlandmarks(), opt->nedges());
	return 0;
}
