// cudns_run -- host driver on top of the C ABI (include/cudns.h): what main.cpp:27-90 + solverWrapper (cuda_main.cu:267-327) do in
// the reference, with run-time configuration instead of compile-time globals.h.  One process; ngpus=N runs N z-slabs on N GPUs of
// the box from N host threads (multi-process runs are driven through cudanavierstokes_b200/dist.py, which owns the rank bootstrap).
//
//   cudns_run [config-file] [key=value ...] [--dry-run]
//
// keys: case=tgv|channel|blayer (presets of python-utils/CompNavierStokes.py, globals/channel.h, src/globals.h), every field of
// cudns_params by name (mx, stencilSize, Re, ...), nsteps, nfiles, restartFile (-1: fresh start), outdir (default "."),
// blasius=internal|<dir with {x,r,u,w,e}Prof.bin>, async_io=0|1, xdmf=0|1, ngpus=N (z-slabs on devices device..device+N-1 of this box; samedevice=1: all on one, for tests), par2_enstrophy=0|1 (Taylor-Green dissipation history in the
// par2 column), post=<first>:<last> (no time stepping: the reference's post-processing tool postproc/post.cpp over the saved
// fields/<c>.<first..last>.bin of outdir -> mean.txt, fluc.txt, bulk.txt there).
// Outputs, in outdir, with the reference's names and formats: Grid.txt, fields/{x,y,z}.bin, fields/{r,u,v,w,e}.<%07d>.bin,
// solution.txt, prof.txt, inProf.txt / inRef.txt (boundary layer) (+ fields.xmf).  --dry-run stops before the GPU is touched (grid, initial condition, file 0).
#include <cudns.h>

#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <map>
#include <mutex>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <vector>

namespace {

struct Field { const char *name; size_t off; char type; };     // 'i' int, 'd' double
#define F_I(n) {#n, offsetof(cudns_params, n), 'i'}
#define F_D(n) {#n, offsetof(cudns_params, n), 'd'}
const Field kFields[] = {
    F_I(mx), F_I(my), F_I(mz), F_I(stencilSize), F_I(stencilVisc), F_D(Lx), F_D(Ly), F_D(Lz), F_D(CFL), F_I(lowStorage),
    F_I(boundaryLayer), F_I(perturbed), F_I(forcing), F_I(periodicX), F_I(nonUniformX), F_I(checkCFLcondition), F_I(checkBulk),
    F_D(Re), F_D(Pr), F_D(Ma), F_D(viscexp), F_D(gam), F_D(stretch), F_D(TwallTop), F_D(TwallBot),
    F_D(spTopStr), F_D(spTopLen), F_D(spTopExp), F_D(spInlStr), F_D(spInlLen), F_D(spInlExp), F_D(spOutStr), F_D(spOutLen), F_D(spOutExp),
    F_I(kC), F_I(LP), F_D(amp1), F_D(amp2), F_D(omega1), F_D(omega2), F_I(quirk_q1), F_I(rk4), F_I(device), F_I(par2_enstrophy), F_I(precision)};

void die(const std::string &msg) { std::fprintf(stderr, "cudns_run: %s\n", msg.c_str()); std::exit(1); }
#define CK(call) do { if ((call) != CUDNS_OK) die(std::string(#call) + ": " + cudns_last_error()); } while (0)

void parse_kv(const std::string &tok, std::map<std::string, std::string> &kv) {
    const size_t eq = tok.find('=');
    if (eq == std::string::npos || eq == 0) die("expected key=value, got '" + tok + "'");
    kv[tok.substr(0, eq)] = tok.substr(eq + 1);
}

void read_config(const char *path, std::map<std::string, std::string> &kv) {
    FILE *f = std::fopen(path, "r");
    if (!f) die(std::string("cannot open config ") + path);
    char line[1024];
    while (std::fgets(line, sizeof(line), f)) {
        std::string s(line);
        const size_t hash = s.find('#');
        if (hash != std::string::npos) s.erase(hash);
        std::string t;
        for (char c : s) if (c != ' ' && c != '\t' && c != '\n' && c != '\r') t.push_back(c);
        if (!t.empty()) parse_kv(t, kv);
    }
    std::fclose(f);
}

std::vector<double> read_bin(const std::string &path, size_t n) {
    std::vector<double> v(n);
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) die("cannot open " + path);
    if (std::fread(v.data(), sizeof(double), n, f) != n) die("short read " + path);
    std::fclose(f);
    return v;
}

void write_bin(const std::string &path, const double *v, size_t n) {
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) die("cannot open " + path);
    if (std::fwrite(v, sizeof(double), n, f) != n) die("short write " + path);
    std::fclose(f);
}


// ---- the slabs of a multi-GPU run inside one process: one host thread per GPU (the library's cudns_team wires the solvers together)
struct Team {
    int n;
    std::mutex m; std::condition_variable cv; int waiting = 0; unsigned long gen = 0; uint64_t written = 0;
    explicit Team(int n_) : n(n_) {}
    void barrier() {
        std::unique_lock<std::mutex> lk(m);
        const unsigned long g = gen;
        if (++waiting == n) { waiting = 0; gen++; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
    void add_written(uint64_t k) { std::lock_guard<std::mutex> lk(m); written += k; }
    template <class F> void run(F &&f) {
        if (n == 1) { f(0); return; }
        std::vector<std::thread> th;
        for (int r = 0; r < n; r++) th.emplace_back([&, r] { f(r); });
        for (auto &t : th) t.join();
    }
};

}  // namespace

int main(int argc, char **argv) {
    std::map<std::string, std::string> kv;
    bool dry = false;
    for (int a = 1; a < argc; a++) {
        const std::string tok = argv[a];
        if (tok == "--dry-run") dry = true;
        else if (tok.find('=') == std::string::npos) read_config(argv[a], kv);
        else parse_kv(tok, kv);
    }
    auto take = [&](const char *key, const std::string &def) { auto it = kv.find(key); if (it == kv.end()) return def; std::string v = it->second; kv.erase(it); return v; };
    const std::string cas = take("case", "tgv");
    const int nsteps = std::atoi(take("nsteps", "100").c_str()), nfiles = std::atoi(take("nfiles", "1").c_str());
    const int restartFile = std::atoi(take("restartFile", "-1").c_str());
    const std::string outdir = take("outdir", "."), blasius = take("blasius", "internal");
    const bool async_io = std::atoi(take("async_io", "1").c_str()) != 0, xdmf = std::atoi(take("xdmf", "1").c_str()) != 0;
    const std::string post = take("post", "");
    const int ngpus = std::atoi(take("ngpus", "1").c_str());
    const bool samedevice = std::atoi(take("samedevice", "0").c_str()) != 0;      // testing aid: all slabs on ONE device
    if (ngpus < 1 || ngpus > 64) die("ngpus must be in 1..64");
    if (nsteps < 2 || nfiles < 1) die("nsteps must be >= 2 and nfiles >= 1");

    cudns_params P;
    if (cas == "tgv") {
        int n = 64, st = 3;
        if (kv.count("mx")) n = std::atoi(kv["mx"].c_str());
        if (kv.count("stencilSize")) st = std::atoi(kv["stencilSize"].c_str());
        CK(cudns_params_tgv(&P, n, st));
    } else if (cas == "channel") CK(cudns_params_channel(&P));
    else if (cas == "blayer") CK(cudns_params_blayer(&P));
    else die("case must be tgv, channel or blayer");
    for (const auto &it : kv) {
        const Field *fd = nullptr;
        for (const Field &f : kFields) if (it.first == f.name) fd = &f;
        if (!fd) die("unknown key '" + it.first + "'");
        if (fd->type == 'i') *(int *)((char *)&P + fd->off) = std::atoi(it.second.c_str());
        else *(double *)((char *)&P + fd->off) = std::atof(it.second.c_str());
    }
    if (kv.count("stencilSize") && !kv.count("stencilVisc") && cas == "tgv") P.stencilVisc = P.stencilSize;
    if (!kv.count("omega1")) P.omega1 = P.Re * 121.e-6;                  // perturbation.h:20 derives it from Re
    P.nranks = 1; P.rank = 0;                                            // (host set-up below works on the global grid)

    // ---- initGrid (init.cpp:32-92): grid, Grid.txt, fields/{x,y,z}.bin
    ::mkdir(outdir.c_str(), 0755); ::mkdir((outdir + "/fields").c_str(), 0755);
    const size_t N = (size_t)P.mx * P.my * P.mz;
    std::vector<double> x(P.mx), xp(P.mx), xpp(P.mx), y(P.my), z(P.mz);
    double dx = 0.0;
    CK(cudns_init_grid(&P, x.data(), xp.data(), xpp.data(), y.data(), z.data(), &dx));
    {
        FILE *fp = std::fopen((outdir + "/Grid.txt").c_str(), "w+");
        if (!fp) die("cannot write Grid.txt");
        for (int i = 0; i < P.mx; i++) std::fprintf(fp, "%d %lf %lf %lf\n", i, x[i], xp[i], xpp[i]);
        std::fclose(fp);
        write_bin(outdir + "/fields/x.bin", x.data(), x.size());
        write_bin(outdir + "/fields/y.bin", y.data(), y.size());
        write_bin(outdir + "/fields/z.bin", z.data(), z.size());
    }
    // ---- calculateSponge (sponge.cu:83-240) / restartWrapper: sponge tables and the initial condition
    std::vector<double> r, u, v, w, e, sigx, sigz, ref5;
    const bool fresh = restartFile < 0;
    if (fresh) { r.resize(N); u.resize(N); v.resize(N); w.resize(N); e.resize(N); }
    if (P.boundaryLayer) {
        const int n = 1000;
        std::vector<double> bx(n), br(n), bu(n), bw(n), be(n);
        if (blasius == "internal") CK(cudns_blasius_profiles(P.gam, P.Ma, P.Pr, n, bx.data(), br.data(), bu.data(), bw.data(), be.data()));
        else {
            bx = read_bin(blasius + "/xProf.bin", n); br = read_bin(blasius + "/rProf.bin", n); bu = read_bin(blasius + "/uProf.bin", n);
            bw = read_bin(blasius + "/wProf.bin", n); be = read_bin(blasius + "/eProf.bin", n);
        }
        sigx.resize(P.mx); sigz.resize(P.mz); ref5.resize(5 * (size_t)P.mx * P.mz);
        // the reference's 1-based spline skips the first knot (quirk Q9): same convention as the parity tests
        CK(cudns_build_sponge(&P, x.data(), z.data(), bx.data() + 1, br.data() + 1, bu.data() + 1, bw.data() + 1, n - 1, sigx.data(), sigz.data(),
                              ref5.data(), fresh ? r.data() : nullptr, fresh ? u.data() : nullptr, fresh ? v.data() : nullptr,
                              fresh ? w.data() : nullptr, fresh ? e.data() : nullptr));
        {   // the two text files calculateSponge leaves behind (sponge.cu:178-182,196-200): the similarity profiles as read, and the
            // inflow reference state (plane k = 0 of the reference tables = of a fresh initial field)
            FILE *fw = std::fopen((outdir + "/inProf.txt").c_str(), "w+");
            if (!fw) die("cannot write inProf.txt");
            for (int i = 0; i < n; i++) std::fprintf(fw, "%le %le %le %le %le\n", bx[i], br[i], bu[i], bw[i], be[i]);
            std::fclose(fw);
            fw = std::fopen((outdir + "/inRef.txt").c_str(), "w+");
            if (!fw) die("cannot write inRef.txt");
            const size_t nt = (size_t)P.mx * P.mz;                       // ref5 = (rho, rho u, rho v, rho w, rho E) tables [mx*mz], index i + k*mx
            for (int i = 0; i < P.mx; i++) {
                const double rr = ref5[i];
                std::fprintf(fw, "%le %le %le %le %le\n", x[i], rr, ref5[nt + i] / rr, ref5[3 * nt + i] / rr, ref5[4 * nt + i]);
            }
            std::fclose(fw);
        }
    } else if (fresh) {
        if (P.forcing) CK(cudns_init_channel(&P, x.data(), y.data(), z.data(), r.data(), u.data(), v.data(), w.data(), e.data()));
        else CK(cudns_init_chit(&P, x.data(), y.data(), z.data(), r.data(), u.data(), v.data(), w.data(), e.data()));
    }
    std::printf("cudns_run: %s  case %s  grid %d x %d x %d  s=%d v=%d  %s  nfiles %d x nsteps %d  outdir %s\n", cudns_version(), cas.c_str(),
                P.mx, P.my, P.mz, P.stencilSize, P.stencilVisc, P.rk4 ? "RK4" : P.lowStorage ? "low-storage RK3" : "Kutta RK3", nfiles, nsteps, outdir.c_str());
    if (dry) {
        if (fresh) {
            const double *fl[5] = {r.data(), u.data(), v.data(), w.data(), e.data()};
            const char nm[5] = {'r', 'u', 'v', 'w', 'e'};
            for (int f = 0; f < 5; f++) CK(cudns_write_field(outdir.c_str(), nm[f], 0, fl[f], N));
        }
        const int ts0 = 0;
        if (xdmf) CK(cudns_write_xdmf((outdir + "/fields/fields.xmf").c_str(), 0, x.data(), P.mx, y.data(), P.my, z.data(), P.mz, &ts0, 1, 0.0, "ruvwe"));
        std::printf("cudns_run: dry run, stopping before the GPU is touched\n");
        return 0;
    }

    // ---- setDevice + setGPUParameters + initSolver + copyField(0) (main.cpp:62-66), one solver per GPU.  ngpus > 1: z-slabs inside
    // this one process (one host thread per GPU, the role of the reference's MPI ranks): the slabs' state blocks are mapped into each
    // other (cudns_halo_connect, same-process branch: peer access) so that the stage kernel stores its boundary planes straight into the
    // neighbours' ghost planes over NVLink; scalar reductions and the one-off ghost fill of copyField(0) go through the two callbacks below
    P.nranks = ngpus;
    if (P.mz % ngpus) die("mz must be divisible by ngpus");
    const int mzl = P.mz / ngpus;
    const size_t Nl = (size_t)P.mx * P.my * mzl;
    Team team(ngpus);
    std::vector<cudns_handle> H(ngpus, nullptr);
    for (int rk = 0; rk < ngpus; rk++) {
        cudns_params Pr = P; Pr.rank = rk; Pr.device = samedevice ? P.device : P.device + rk;
        CK(cudns_create(&Pr, x.data(), xp.data(), xpp.data(), &H[rk]));
    }
    cudns_team_handle group = nullptr;
    if (ngpus > 1) CK(cudns_team_create(H.data(), ngpus, &group));
    if (!post.empty()) {               // postproc/post.cpp:126-199 instead of the time loop
        int first = 0, last = 0;
        if (std::sscanf(post.c_str(), "%d:%d", &first, &last) != 2) die("post=<first>:<last> expected");
        std::printf("Fields to postprocess :  %d -> %d\n", first, last);
        team.run([&](int rk) { CK(cudns_postprocess(H[rk], outdir.c_str(), first, last, x.data(), outdir.c_str())); });
        for (int rk = 0; rk < ngpus; rk++) CK(cudns_destroy(H[rk]));
        if (group) CK(cudns_team_destroy(group));
        std::printf("cudns_run: wrote mean.txt, fluc.txt, bulk.txt\n");
        return 0;
    }
    // this rank's slab of the sponge tables: ref5 is [5][mz][mx], sigma_z is [mz]
    auto set_sponge = [&](int rk) {
        if (!P.boundaryLayer) return;
        std::vector<double> rf(5 * (size_t)P.mx * mzl);
        for (int f = 0; f < 5; f++)
            std::memcpy(rf.data() + (size_t)f * P.mx * mzl, ref5.data() + (size_t)f * P.mx * P.mz + (size_t)rk * P.mx * mzl, sizeof(double) * P.mx * mzl);
        CK(cudns_set_sponge(H[rk], sigx.data(), sigz.data() + (size_t)rk * mzl, rf.data()));
    };
    auto write_prof = [&](int rk) {        // calcAvgChan (init.cpp:150-208): prof.txt is rewritten at every output (a collective call)
        std::vector<double> prof(10 * (size_t)P.mx);
        CK(cudns_calc_profiles(H[rk], prof.data()));
        if (rk != 0) return;
        FILE *fp = std::fopen((outdir + "/prof.txt").c_str(), "w+");
        if (!fp) die("cannot write prof.txt");
        for (int i = 0; i < P.mx; i++) {
            std::fprintf(fp, "%lf", x[i]);
            for (int q = 0; q < 10; q++) std::fprintf(fp, "\t%lf", prof[(size_t)q * P.mx + i]);
            std::fprintf(fp, "\n");
        }
        std::fclose(fp);
    };
    std::vector<int> saved;
    const int start = fresh ? 0 : restartFile;
    FILE *sol = std::fopen((outdir + "/solution.txt").c_str(), "w+");
    if (!sol) die("cannot write solution.txt");
    std::chrono::steady_clock::time_point t0;
    uint64_t nwritten = 0;
    team.run([&](int rk) {
        set_sponge(rk);
        if (fresh) CK(cudns_set_state(H[rk], r.data() + rk * Nl, u.data() + rk * Nl, v.data() + rk * Nl, w.data() + rk * Nl, e.data() + rk * Nl));
        else CK(cudns_read_fields(H[rk], outdir.c_str(), restartFile));
        {   // calcdt + printRes + calcAvgChan of the initial field (main.cpp:57-60)
            double dt0 = 0.0, ret = 0.0;
            CK(cudns_calc_dt(H[rk], &dt0));
            if (rk == 0) std::printf("the initial dt is : %lf\n", dt0);
            if (!P.periodicX) { CK(cudns_calc_retau(H[rk], &ret)); if (rk == 0) std::printf("The average friction Reynolds number is: \t %lf\n", ret); }
            write_prof(rk);
        }
        if (fresh) { CK(cudns_write_fields_async(H[rk], outdir.c_str(), 0)); if (!async_io) CK(cudns_io_wait(H[rk], nullptr)); if (rk == 0) saved.push_back(0); }
        // ---- solverWrapper (cuda_main.cu:267-327)
        std::vector<double> htime(nsteps), hpar1(nsteps), hpar2(nsteps);
        team.barrier();
        if (rk == 0) t0 = std::chrono::steady_clock::now();
        for (int file = start + 1; file < nfiles + start + 1; file++) {
            std::fill(hpar1.begin(), hpar1.end(), 0.0); std::fill(hpar2.begin(), hpar2.end(), 0.0);
            CK(cudns_advance(H[rk], nsteps, htime.data(), hpar1.data(), hpar2.data()));
            CK(cudns_write_fields_async(H[rk], outdir.c_str(), file));         // copyField(1) + writeField(file) without stalling the next file
            if (!async_io) CK(cudns_io_wait(H[rk], nullptr));
            write_prof(rk);
            if (rk != 0) continue;
            saved.push_back(file);
            // par1 / par2 are refreshed every checkBulk steps: report the last refreshed entry like the reference's device arrays hold it
            int last = ((nsteps - 1) / P.checkBulk) * P.checkBulk;
            std::printf("file number: %d  \t step: %d  \t time: %lf  \t kin: %le  \t energy: %le\n", file, file * nsteps, htime[nsteps - 1], hpar1[last], hpar2[last]);
            for (int t = 0; t < nsteps - 1; t += P.checkCFLcondition)
                std::fprintf(sol, "%d %lf %lf %lf %lf\n", file * (t + 1), htime[t], hpar1[t], hpar2[t], htime[t + 1] - htime[t]);
            std::fflush(sol);
        }
        uint64_t nw = 0;
        CK(cudns_io_wait(H[rk], &nw));
        team.add_written(nw);
    });
    std::fclose(sol);
    nwritten = team.written;
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::vector<double>().swap(r); std::vector<double>().swap(u); std::vector<double>().swap(v); std::vector<double>().swap(w); std::vector<double>().swap(e);
    if (xdmf) {
        double dtn = 0.0; CK(cudns_get_scalars(H[0], &dtn, nullptr, nullptr));
        CK(cudns_write_xdmf((outdir + "/fields/fields.xmf").c_str(), 0, x.data(), P.mx, y.data(), P.my, z.data(), P.mz, saved.data(), (int)saved.size(),
                            dtn * nsteps, "ruvwe"));
    }
    const int stages = P.rk4 ? 4 : 3;
    std::printf("The total time is: %lf\nThe simulation time per time step is: %lf\n", secs, secs / ((double)nfiles * nsteps));
    std::printf("cudns_run: %.1f Mpts*RK-stage/s on %d GPU%s (step loop + diagnostics + output hand-off), %llu field files written\n",
                (double)N * stages * nfiles * nsteps / secs / 1e6, ngpus, ngpus > 1 ? "s" : "", (unsigned long long)nwritten);
    for (int rk = 0; rk < ngpus; rk++) CK(cudns_destroy(H[rk]));
    if (group) CK(cudns_team_destroy(group));
    std::printf("Simulation is finished! \n");
    return 0;
}
