// The qvis3 executable: same command line as the reference tool (src/tools/qvis3/qvis3.c:493-566); -fast only.
#include "../../include/qvis_cuda.h"

int main(int argc, char **argv) { return qvis3_main(argc, argv); }
