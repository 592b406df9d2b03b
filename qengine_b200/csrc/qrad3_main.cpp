// The qrad3 executable: same command line as the reference tool (src/tools/qrad3/qrad3.c:545-643).
#include "../../include/qrad_host.h"

int main(int argc, char **argv) { return qrad3_main(argc, argv); }
