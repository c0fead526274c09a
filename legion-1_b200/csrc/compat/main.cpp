// main.cpp -- `legion <ngpu> <cache_agg_mode>`: the reference's entry point (main.cpp:4-10), unchanged in shape.
#include <cstdlib>
#include <cstdio>

#include "Server.h"

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: legion <gpu number> <cache aggregate mode 0|1|2|3>  (run in the directory holding ./meta_config)\n"); return 2; }
    Server* server = NewGPUServer();
    server->Initialize(atoi(argv[1]));   // gpu number
    server->PreSc(atoi(argv[2]));        // cache aggregate mode: GPUs per NVLink clique = 1 / 2 / 4 / 8
    server->Run();
    server->Finalize();
    return 0;
}
