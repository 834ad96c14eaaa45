// index_assignment_driver.cpp -- TEST DRIVER (needs no GPU): initialize() on a .cubagraph, then dumps the internal indices the class
// wrote into the caller's vertices (iP per pose, iL per landmark, file order) as int32.  tests/test_host_initialize.py compares them
// with graphio.flatten's numbering (reference src/cuda_bundle_adjustment.cpp:142-200: ascending id, free first, fixed appended,
// vertices without edges skipped).
#include "../../samples/cubagraph_reader.h"

int main(int argc, char** argv)
{
	if (argc < 3) { fprintf(stderr, "usage: index_assignment_driver graph.cubagraph out.bin [repeats]\n"); return 2; }
	Storage st;
	auto opt = readGraph(argv[1], st);
	const int reps = argc > 3 ? atoi(argv[3]) : 1;
	for (int i = 0; i < reps; i++) opt->initialize();      // repeated calls reuse the cached order and the list pool
	FILE* f = fopen(argv[2], "wb");
	if (!f) return 2;
	for (auto& v : st.poses) { const int32_t x = v->iP; fwrite(&x, 4, 1, f); }
	for (auto& v : st.landmarks) { const int32_t x = v->iL; fwrite(&x, 4, 1, f); }
	fclose(f);
	printf("{\"nposes\": %zu, \"nlandmarks\": %zu, \"nedges\": %zu}\n", opt->nposes(), opt->nlandmarks(), opt->nedges());
	return 0;
}
