// kcftools_b200 — command line front end of the B200 getVariations drop-in (see kcf_host.hpp).
#include "kcf_host.hpp"

int main(int argc, char **argv) { return kcfh::cliMain(argc, argv); }
