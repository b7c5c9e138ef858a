/* Test-only driver for bt_build.h / bt_build_sa.cuh over the host backend (bsa_host.h): the same templates and per-element
 * functors the product runs on the GPU, executed sequentially with std:: algorithms in place of CUB. */
#include "bsa_host.h"

int main(int argc, char **argv) {
	BtBuildParams P; std::vector<std::string> fa; std::string base;
	for (int i = 1; i < argc; i++) {
		std::string a = argv[i];
		if (a == "-o") P.offRate = atoi(argv[++i]);
		else if (a == "-t") P.ftabChars = atoi(argv[++i]);
		else if (a == "--ntoa") P.nsToAs = true;
		else if (fa.empty()) { size_t p0 = 0; for (;;) { size_t c = a.find(',', p0); fa.push_back(a.substr(p0, c == std::string::npos ? c : c - p0)); if (c == std::string::npos) break; p0 = c + 1; } }
		else base = a;
	}
	std::string err;
	if (!host_build_all(fa, base, P, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
	return 0;
}
