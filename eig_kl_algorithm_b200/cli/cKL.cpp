// cKL -- drop-in for the reference's cKL executable (cKL.cpp:424-468), GPU backed.
#include "kl_main.h"
int main(int argc, char *argv[]) { return kl_main(argc, argv, false); }
