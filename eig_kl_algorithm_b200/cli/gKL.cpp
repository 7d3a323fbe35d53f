// gKL -- drop-in for the reference's gKL executable (gKL.cu:672-713).  Same engine as cKL (the
// reference's gKL differs from cKL only in float summation order and never writes its trace file,
// gKL.cu:689-690 -- this one does, in cKL's format); argc < 2 prints the usage on stderr.
#include "kl_main.h"
int main(int argc, char *argv[]) { return kl_main(argc, argv, true); }
