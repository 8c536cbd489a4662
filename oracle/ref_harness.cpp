// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Thin driver around the UNMODIFIED reference engine (class `family`,
// /root/reference/src/family.{h,cpp}).  It is compiled by oracle/Makefile from the
// reference sources where they lie (nothing is copied into this repo) into
// oracle/_ref/ref_harness.  It feeds raw FP64 likelihood batches through the same five
// calls the reference drivers make per variant (src/file.cpp:595-682, :1743-1806):
//   set_LK -> calPostProb{BN,Peeling,MCMC} -> get_postProb / get_postProbSingle / get_postRlt
// and dumps the raw doubles, so parity is checked on numbers, not on "%g" text.
//
// usage: ref_harness key=value ... in=<batch.bin> out=<result.bin>
//   ped=<file>  method=1|2|3  mrate=<x>  lc=<x>  burn=<n>  rep=<n>  seed=<n|-1>
//   gpn=a,b,c gpk=a,b,c gpxn=a,b,c gpxk=a,b,c   (optional prior overrides, 3 values each)
//   cols=i,j,k   ped-row index of every sequenced input column, in input-column order
//   repeat=<r>   run the whole batch r times (timing); results are from the last pass
//   limit=<v>    only the first v variants
// batch.bin : int32 V, int32 S, uint8 flags[V] (bit0 Known, bit1 chrX), double lk[V][S][3]
// result.bin: int32 V, int32 S, int32 N, uint8 status[V] (1 = engine returned false),
//             double post[V][S][3], double single[V][S][3], int32 gt[V][S],
//             double post_full[V][N][3], double single_full[V][N][3]
// stdout    : one JSON line with the engine-only elapsed seconds.

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "family.h"
#include "file.h"

static std::vector<double> parse_list(const std::string &s) {
    std::vector<double> v;
    std::stringstream ss(s);
    std::string tok;
    while (std::getline(ss, tok, ',')) v.push_back(atof(tok.c_str()));
    return v;
}

int main(int argc, char **argv) {
    std::map<std::string, std::string> kv;
    for (int i = 1; i < argc; i++) {
        std::string a(argv[i]);
        size_t p = a.find('=');
        if (p == std::string::npos) { fprintf(stderr, "bad arg %s\n", argv[i]); return 2; }
        kv[a.substr(0, p)] = a.substr(p + 1);
    }
    auto get = [&](const char *k, const char *d) { return kv.count(k) ? kv[k] : std::string(d); };
    int method = atoi(get("method", "1").c_str());
    double mrate = atof(get("mrate", "1e-7").c_str());
    double lc = atof(get("lc", "1").c_str());
    int burn = atoi(get("burn", "1000").c_str());
    int rep = atoi(get("rep", "100000").c_str());
    long seed = atol(get("seed", "-1").c_str());
    int repeat = atoi(get("repeat", "1").c_str());
    long limit = atol(get("limit", "-1").c_str());

    std::vector<individual> mem;
    if (!readPed(get("ped", ""), mem)) return 3;
    family fam(mem, mrate);
    if (kv.count("gpn")) fam.set_genoProbN(parse_list(kv["gpn"]));
    if (kv.count("gpk")) fam.set_genoProbK(parse_list(kv["gpk"]));
    if (kv.count("gpxn")) fam.set_genoProbXN(parse_list(kv["gpxn"]));
    if (kv.count("gpxk")) fam.set_genoProbXK(parse_list(kv["gpxk"]));
    fam.set_lc(lc);
    if (!fam.init()) { fprintf(stderr, "family::init failed\n"); return 4; }
    const int N = (int)fam.get_numInd();

    std::vector<double> colsd = parse_list(get("cols", ""));
    std::vector<int> mapV2P, mapP2V(N, -1);
    for (size_t i = 0; i < colsd.size(); i++) {
        mapV2P.push_back((int)colsd[i]);
        mapP2V[(int)colsd[i]] = (int)i;
    }
    fam.set_mapP2V(mapP2V);
    fam.set_mapV2P(mapV2P);

    FILE *fi = fopen(get("in", "").c_str(), "rb");
    if (!fi) { fprintf(stderr, "cannot open input batch\n"); return 5; }
    int V = 0, S = 0;
    if (fread(&V, 4, 1, fi) != 1 || fread(&S, 4, 1, fi) != 1) return 5;
    if (S != (int)mapV2P.size()) { fprintf(stderr, "S mismatch\n"); return 5; }
    std::vector<unsigned char> flags(V);
    std::vector<double> lk((size_t)V * S * 3);
    if (V && fread(flags.data(), 1, V, fi) != (size_t)V) return 5;
    if (V && fread(lk.data(), 8, lk.size(), fi) != lk.size()) return 5;
    fclose(fi);
    if (limit >= 0 && limit < V) V = (int)limit;

    std::vector<unsigned char> status(V, 0);
    std::vector<double> post((size_t)V * S * 3, 0), single((size_t)V * S * 3, 0);
    std::vector<int> gt((size_t)V * S, -1);
    std::vector<double> postF((size_t)V * N * 3, 0), singleF((size_t)V * N * 3, 0);

    if (seed >= 0) srand((unsigned)seed);
    double elapsed = 0;
    for (int r = 0; r < repeat; r++) {
        auto t0 = std::chrono::steady_clock::now();
        for (int v = 0; v < V; v++) {
            dMatrix<double> LK(N, 3, 1);
            for (int s = 0; s < S; s++)
                for (int g = 0; g < 3; g++) LK(mapV2P[s], g) = lk[((size_t)v * S + s) * 3 + g];
            fam.set_LK(LK);
            bool known = flags[v] & 1;
            int chrType = (flags[v] >> 1) & 1;
            bool ok = false;
            if (method == 1) ok = fam.calPostProbBN(known, chrType);
            else if (method == 2) ok = fam.calPostProbPeeling(known, chrType);
            else ok = fam.calPostProbMCMC(burn, rep, known, chrType);
            status[v] = ok ? 0 : 1;
            if (!ok) continue;
            dMatrix<double> pp = fam.get_postProb();
            std::vector<int> rl = fam.get_postRlt();
            dMatrix<double> ps = fam.get_postProbSingle();
            dMatrix<double> ppF = fam.get_postProb(false);
            dMatrix<double> psF = fam.get_postProbSingle(false);
            for (int s = 0; s < S; s++) {
                for (int g = 0; g < 3; g++) {
                    post[((size_t)v * S + s) * 3 + g] = pp(s, g);
                    single[((size_t)v * S + s) * 3 + g] = ps(s, g);
                }
                gt[(size_t)v * S + s] = rl[s];
            }
            for (int i = 0; i < N; i++)
                for (int g = 0; g < 3; g++) {
                    postF[((size_t)v * N + i) * 3 + g] = ppF(i, g);
                    singleF[((size_t)v * N + i) * 3 + g] = psF(i, g);
                }
        }
        auto t1 = std::chrono::steady_clock::now();
        elapsed += std::chrono::duration<double>(t1 - t0).count();
    }

    if (kv.count("out")) {
        FILE *fo = fopen(kv["out"].c_str(), "wb");
        if (!fo) return 6;
        fwrite(&V, 4, 1, fo);
        fwrite(&S, 4, 1, fo);
        fwrite(&N, 4, 1, fo);
        fwrite(status.data(), 1, V, fo);
        fwrite(post.data(), 8, (size_t)V * S * 3, fo);
        fwrite(single.data(), 8, (size_t)V * S * 3, fo);
        fwrite(gt.data(), 4, (size_t)V * S, fo);
        fwrite(postF.data(), 8, (size_t)V * N * 3, fo);
        fwrite(singleF.data(), 8, (size_t)V * N * 3, fo);
        fclose(fo);
    }
    printf("{\"variants\": %d, \"repeat\": %d, \"elapsed_s\": %.6f, \"variants_per_s\": %.3f}\n", V, repeat,
           elapsed, elapsed > 0 ? (double)V * repeat / elapsed : 0.0);
    return 0;
}
