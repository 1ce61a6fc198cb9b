// no-op stages of the reference (submission/src/bin/client_preprocess.rs, server_preprocess_dataset.rs)
#include "stage_common.h"
int main(int argc, char **argv) { long size; return parse_size(argc, argv, &size) ? 0 : 1; }
