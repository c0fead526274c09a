// Entry point of the stand-alone sampling server.  Command line and phase order are the reference's
// (reference main.cpp:4-10): `legion <gpu number> <cache aggregate mode>`, run from the directory that holds
// ./meta_config; phases Initialize -> PreSc -> Run -> Finalize on one Server object.
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <memory>

#include "Server.h"

namespace {

// strict decimal parse: the reference atoi()s and silently runs with 0 GPUs on a typo
bool parse_int(const char* text, long lo, long hi, int* out)
{
    char* end = nullptr;
    errno = 0;
    long v = strtol(text, &end, 10);
    if (errno != 0 || end == text || *end != '\0' || v < lo || v > hi) return false;
    *out = (int)v;
    return true;
}

int usage(const char* argv0)
{
    fprintf(stderr,
            "usage: %s <gpu number 1..8> <cache aggregate mode 0|1|2|3>\n"
            "  cache aggregate mode m: GPUs per NVLink clique = 1 << m\n"
            "  reads ./meta_config (written by legion_server.py) from the current directory\n",
            argv0);
    return 2;
}

}  // namespace

int main(int argc, char** argv)
{
    int n_gpus = 0, agg_mode = 0;
    if (argc < 3 || !parse_int(argv[1], 1, 8, &n_gpus) || !parse_int(argv[2], 0, 3, &agg_mode)) return usage(argv[0]);

    std::unique_ptr<Server> server(NewGPUServer());
    server->Initialize(n_gpus);
    server->PreSc(agg_mode);
    server->Run();
    server->Finalize();
    return 0;
}
