// fuzz_host_readers.cpp -- mutation fuzz of the host-side readers behind the drop-in boundary (Ply::Parse +
// GaussianCloud::ImportPly, CamerasConfig::ImportJson), meant to be built with -fsanitize=address,undefined
// (tests/test_host_fuzz.py).  usage: fuzz_host_readers <seed.ply> <seed.json> <scratch dir> [iterations]
// The GPU decode entry points are stubbed: only host code is linked.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <random>
#include <sstream>
#include <string>
#include <vector>
#include "splatapult_b200/camerasconfig.h"
#include "splatapult_b200/gaussiancloud.h"
#include "splatapult_b200/ply.h"
#include "plydecode.h"
namespace splat {
bool decode_ply_on_device(const unsigned char*, size_t, size_t, const PlyDecodeLayout&, bool, int, void*, float*, void**, std::string* e) { if (e) *e = "stub"; return false; }
void free_decoded_aos(void*) {}
}
static std::string slurp(const char* p) { std::ifstream f(p, std::ios::binary); std::stringstream s; s << f.rdbuf(); return s.str(); }
int main(int argc, char** argv)
{
    const std::string ply = slurp(argv[1]);
    const std::string json = slurp(argv[2]);
    std::mt19937 rng(12345);
    const std::string dir = argv[3];
    const std::string tmpPlyS = dir + "/m.ply", tmpJsonS = dir + "/m.json";
    const char* tmpPly = tmpPlyS.c_str();
    const char* tmpJson = tmpJsonS.c_str();
    const int iters = argc > 4 ? std::atoi(argv[4]) : 20000;
    const char* tokens[] = {"property", "float", "element", "vertex", "end_header", "format", "binary_little_endian", "1.0", "uchar", "double", "list", "int", "\n", " ", "999999999999", "-1", "0", "ply"};
    int okPly = 0, okJson = 0;
    for (int it = 0; it < iters; it++)
    {
        std::string m = ply;
        const size_t hdrEnd = m.find("end_header");
        const int nmut = 1 + rng() % 4;
        for (int k = 0; k < nmut; k++)
        {
            const int kind = rng() % 5;
            const size_t lim = (rng() % 4) ? hdrEnd + 11 : m.size();
            if (m.empty()) break;
            const size_t pos = rng() % (lim ? std::min(lim, m.size()) : 1);
            if (kind == 0) m[pos] = (char)(rng() & 0xff);
            else if (kind == 1) m.erase(pos, 1 + rng() % 8);
            else if (kind == 2) m.insert(pos, tokens[rng() % 18]);
            else if (kind == 3) m.resize(pos);
            else m.insert(pos, std::string(1 + rng() % 3, (char)('0' + rng() % 10)));
        }
        { std::ofstream f(tmpPly, std::ios::binary); f.write(m.data(), (std::streamsize)m.size()); }
        GaussianCloud gc(GaussianCloud::Options{true, true});
        if (gc.ImportPly(tmpPly)) { okPly++; (void)gc.GetTotalSize(); }
        std::string j = json;
        const int jm = 1 + rng() % 3;
        for (int k = 0; k < jm && !j.empty(); k++)
        {
            const size_t pos = rng() % j.size();
            const int kind = rng() % 4;
            if (kind == 0) j[pos] = (char)(rng() & 0x7f);
            else if (kind == 1) j.erase(pos, 1 + rng() % 6);
            else if (kind == 2) j.insert(pos, (rng() & 1) ? "[" : "{\"a\":");
            else j.resize(pos);
        }
        { std::ofstream f(tmpJson, std::ios::binary); f.write(j.data(), (std::streamsize)j.size()); }
        CamerasConfig cc;
        if (cc.ImportJson(tmpJson)) { okJson++; (void)cc.GetNumCameras(); }
    }
    // fixed corpus: a header without property lines claiming 2^62 vertices (vertex size 0: the size guard
    // must still reject it), and a cameras.json nested 200 000 levels deep (the recursive parser must refuse, not overflow)
    {
        { std::ofstream f(tmpPly, std::ios::binary); f << "ply\nformat binary_little_endian 1.0\nelement vertex 4611686018427387904\nend_header\n"; }
        GaussianCloud gc(GaussianCloud::Options{true, true});
        if (gc.ImportPly(tmpPly)) { std::printf("corpus: huge no-property ply was accepted\n"); return 1; }
        { std::ofstream f(tmpJson, std::ios::binary); f << std::string(200000, '['); }
        CamerasConfig cc;
        if (cc.ImportJson(tmpJson)) { std::printf("corpus: deep json was accepted\n"); return 1; }
        { std::ofstream f(tmpJson, std::ios::binary); for (int i = 0; i < 100000; i++) f << "{\"a\":"; }
        if (cc.ImportJson(tmpJson)) { std::printf("corpus: deep json object was accepted\n"); return 1; }
    }
    std::printf("fuzz done: ply accepted %d, json accepted %d of %d\n", okPly, okJson, iters);
    return 0;
}
