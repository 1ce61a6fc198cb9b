#include "client_decrypt_common.h"
int main(int argc, char **argv) { return postprocess_stage(argc, argv, "decoded_result.txt", "result.txt"); }
