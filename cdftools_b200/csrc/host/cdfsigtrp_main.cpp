// cdfsigtrp_gpu -- C++ twin of the cdfsigtrp command line (src/cdfsigtrp.f90) on top of libcdfgpu.
// Same required arguments (-t -u -v -smin -smax -nbins) and options -s, -full, -refdep, -neutral, -section, -temp, -teos10,
// -help; the same section file (pairs of lines: name [suffix] [long-name prefix] / imin imax jmin jmax, closed by EOF), the
// same outputs (trpsig.txt, <section>_trpsig.nc | _trptemp.nc) and exit codes (99 usage / missing file, 98 NetCDF, 97 GPU
// library).  The section slices are read and prepared here as at :404-555; the loop nests :559-627 run on the device
// (cdfsigtrp_gpu_section).  Not in this twin: -brk (broken-line files; the library call takes their depths), -xtra, -print,
// -vvl -- refused with a message.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <memory>

#include "host_common.hpp"

using namespace cdfhost;

struct Section {
    std::string name, varname = "none", longname = "none";
    int imin = 0, imax = 0, jmin = 0, jmax = 0;
};

// section_init (:680-748): name lines are cut at single blanks into at most three words; a line holding EOF ends the list
static std::vector<Section> section_init(const std::string &path)
{
    std::ifstream f(path);
    std::vector<Section> out;
    std::string line;
    while (std::getline(f, line)) {
        if (line.find("EOF") != std::string::npos) break;
        Section s;
        std::string w[3] = {"none", "none", "none"};
        int n = 0;
        std::string rest = line + " ";
        size_t ipos = rest.find(' ');
        while (ipos != std::string::npos && ipos >= 1 && n < 3) {   // INDEX(cline,' ') > 1
            w[n++] = rest.substr(0, ipos);
            rest = rest.substr(ipos + 1);
            ipos = rest.find(' ');
        }
        s.name = w[0]; s.varname = w[1]; s.longname = w[2];
        if (!std::getline(f, line)) { printf(" ERROR : %s : section %s has no position line\n", path.c_str(), s.name.c_str()); stop(99); }
        if (sscanf(line.c_str(), "%d %d %d %d", &s.imin, &s.imax, &s.jmin, &s.jmax) != 4) {
            printf(" ERROR : %s : cannot read imin imax jmin jmax of section %s\n", path.c_str(), s.name.c_str());
            stop(99);
        }
        out.push_back(s);
    }
    return out;
}

struct Field {   // one (t,z,y,x) variable of an open file
    nc3::Reader *nc = nullptr;
    const nc3::Var *v = nullptr;
    int nx = 0, ny = 0;
    Field(nc3::Reader &r, const std::string &var, const Names &cn) : nc(&r)
    {
        const int iv = r.find_var(var);
        if (iv < 0) { printf(" ERROR : variable %s not found in %s\n", var.c_str(), r.path.c_str()); stop(98); }
        v = &r.vars[iv];
        nx = (int)r.dim_len(cn.x, true);
        ny = (int)r.dim_len(cn.y, false);
    }
    // getvarxz (cdfio.F90:1950-2043): out[k][0..npts) = var(kimin : kimin+npts-1, kj, k), first time frame
    void xz(int kj, int npts, int npk, int kimin, float *out) const
    {
        for (int k = 0; k < npk; ++k)
            nc_check(nc->read_f32(*v, 0, ((uint64_t)k * ny + (kj - 1)) * nx + (kimin - 1), npts, out + (size_t)k * npts), nc->err);
    }
    // getvaryz (cdfio.F90:2046-2148): out[k][0..npts) = var(ki, kjmin : kjmin+npts-1, k)
    void yz(int ki, int npts, int npk, int kjmin, float *out) const
    {
        for (int k = 0; k < npk; ++k)
            for (int j = 0; j < npts; ++j)
                nc_check(nc->read_f32(*v, 0, ((uint64_t)k * ny + (kjmin - 1 + j)) * nx + (ki - 1), 1, out + (size_t)k * npts + j), nc->err);
    }
};

// Fortran E16.7: 0.ddddddd scaled, exponent of two digits
static std::string e16_7(double x)
{
    char b[64];
    if (x == 0.0 || !std::isfinite(x)) {
        snprintf(b, sizeof b, "%16s", !std::isfinite(x) ? (std::isnan(x) ? "NaN" : (x > 0 ? "Infinity" : "-Infinity")) : "0.0000000E+00");
        return b;
    }
    snprintf(b, sizeof b, "%.6E", fabs(x));   // d.ddddddE+xx, correctly rounded to 7 significant digits
    const int ex = atoi(strchr(b, 'E') + 1) + 1;
    std::string dig;
    dig += b[0];
    dig.append(b + 2, 6);
    char o[64];
    snprintf(o, sizeof o, "%s0.%sE%c%02d", x < 0 ? "-" : "", dig.c_str(), ex < 0 ? '-' : '+', abs(ex));
    snprintf(b, sizeof b, "%16s", o);
    return b;
}

static void file_example()
{
    printf("\n   EXAMPLE of dens_section.dat file\n   --------------------------------\n"
           "  Each section is described by 2 lines :\n"
           "    line#1 : name_of_section [variable_suffix] [ long_name_prefix ]\n"
           "    line#2 : imin  imax jmin jmax \n"
           "     IMPORTANT : the points indicated by imin,jmin  imax,jmax are F-point.\n"
           "  The list is closed by a line holding EOF.\n"
           "  example :\n03_Gibraltar gibra Gibraltar_Strait_transport_in_sigma_classes\n3378 3378 1956 1961\nEOF\n");
}

int main(int argc, char **argv)
{
    Names cn;
    std::string cf_section = "dens_section.dat", cf_out = "trpsig.txt";
    if (argc == 1) {
        printf(" usage :  cdfsigtrp_gpu -t T-file -u U-file -v V-file [-s S-file] -smin sigma_min -smax sigma_max -nbins nbins\n"
               "              [-full] [-refdep ref_depth] [-neutral] [-section file] [-temp] [-help] [-teos10]\n"
               "     PURPOSE : density class transports across the zonal / meridional sections listed in %s,\n"
               "               computed on a B200 GPU.\n"
               "     REQUIRED FILES : %s, %s and %s\n"
               "     OUTPUT : <section>_trpsig.nc (sigma_class, sigtrp in Sv) per section, and %s\n"
               "     NOT IN THIS TWIN : -brk, -xtra, -print, -vvl\n",
               cf_section.c_str(), cn.fhgr.c_str(), cn.fzgr.c_str(), cf_section.c_str(), cf_out.c_str());
        return 0;
    }
    std::string cf_tfil, cf_sfil = "none", cf_ufil, cf_vfil, cglobal;
    double dsigma_min = 0, dsigma_max = 0;
    int nbins = 0, ireq = 0, nreq = 6;
    float refdep = 0.f;
    bool lfull = false, lntr = false, ltemp = false, lteos10 = false;
    for (int i = 0; i < argc; ++i) cglobal += std::string(i ? " " : "") + argv[i];   // SetGlobalAtt: the command line
    for (int i = 1; i < argc;) {   // :228-256
        const std::string a = argv[i++];
        auto next = [&]() { return std::string(i < argc ? argv[i++] : ""); };
        if (a == "-t") { cf_tfil = next(); ++ireq; }
        else if (a == "-u") { cf_ufil = next(); ++ireq; }
        else if (a == "-v") { cf_vfil = next(); ++ireq; }
        else if (a == "-smin") { dsigma_min = atof(next().c_str()); ++ireq; }
        else if (a == "-smax") { dsigma_max = atof(next().c_str()); ++ireq; }
        else if (a == "-nbins") { nbins = atoi(next().c_str()); ++ireq; }
        else if (a == "-s") { cf_sfil = next(); ++ireq; nreq = 7; }
        else if (a == "-full") lfull = true;
        else if (a == "-temp") ltemp = true;
        else if (a == "-help") { file_example(); stop(99); }
        else if (a == "-refdep") refdep = (float)atof(next().c_str());
        else if (a == "-section") cf_section = next();
        else if (a == "-neutral") lntr = true;
        else if (a == "-teos10") lteos10 = true;
        else if (a == "-brk" || a == "-xtra" || a == "-print" || a == "-vvl") {
            printf(" ERROR : %s : not available in the GPU twin (cdfsigtrp_gpu_section takes broken-line depths; the extra\n"
                   "         files and tables are host-side output of the reference).\n", a.c_str());
            stop(99);
        }
        else { printf(" ERROR : %s : unknown option.\n", a.c_str()); stop(99); }
    }
    if (cf_sfil == "none") cf_sfil = cf_tfil;
    if (ireq != nreq) { printf(" ERROR :  not enough input arguments. See the usage message and correct.\n"); stop(99); }
    if (nbins < 1) { printf(" ERROR : -nbins must be positive\n"); stop(99); }
    bool lchk = chkfile(cn.fzgr);
    lchk = chkfile(cn.fhgr) || lchk; lchk = chkfile(cf_section) || lchk; lchk = chkfile(cf_tfil) || lchk;
    lchk = chkfile(cf_sfil) || lchk; lchk = chkfile(cf_ufil) || lchk; lchk = chkfile(cf_vfil) || lchk;
    if (lchk) stop(99);

    nc3::Reader ft, fs, fu, fv, hgr;
    nc_check(ft.open(cf_tfil), ft.err); nc_check(fs.open(cf_sfil), fs.err); nc_check(fu.open(cf_ufil), fu.err);
    nc_check(fv.open(cf_vfil), fv.err); nc_check(hgr.open(cn.fhgr), hgr.err);
    MeshZgr zgr;
    zgr.open(cn.fzgr);
    const Field T(ft, cn.votemper, cn), S(fs, cn.vosaline, cn), U(fu, "vozocrtx", cn), V(fv, cn.vomecrty, cn);
    const float zsps = fs.spval(*S.v, cn.missing), zspu = fu.spval(*U.v, cn.missing), zspv = fv.spval(*V.v, cn.missing);   // :289-291
    int mode = lntr ? 1 : 0;
    if (ltemp) {   // temperature decreases downward: change sign and swap the limits (:303-308)
        mode = 2;
        const double t = dsigma_max;
        dsigma_max = -dsigma_min;
        dsigma_min = -t;
    }
    const std::vector<Section> sec = section_init(cf_section);
    if (sec.empty()) { printf(" ERROR : no section in %s\n", cf_section.c_str()); stop(99); }
    const int npk = (int)ft.dim_len(cn.z, false);
    if (npk < 1) { printf(" ERROR : no vertical dimension %s in %s\n", cn.z.c_str(), cf_tfil.c_str()); stop(98); }
    std::vector<float> gdept(npk), gdepw(npk), e3t1d(npk), e3w1d(npk);
    read_1d(zgr.nc, zgr.name1d("gdept"), npk, gdept.data());
    read_1d(zgr.nc, zgr.name1d("gdepw"), npk, gdepw.data());
    if (lfull) {
        read_1d(zgr.nc, zgr.name1d("e3t1d"), npk, e3t1d.data());
        read_1d(zgr.nc, zgr.ver == "v2.0" ? "e3w" : zgr.ver == "v3.0" ? "e3w_0" : "e3w_1d", npk, e3w1d.data());   // getvare3 (cdfio.F90:2265-2274)
    }
    const std::string sfx = zgr.ver == "v2.0" ? "_ps" : zgr.ver == "v3.0" ? "" : "_0";   // :418-422 / :510-514
    gpu_check(cdfgpu_init(-1, 1), "cdfgpu_init");

    std::vector<double> dsigma_lev(nbins + 1), dtrpbin((size_t)sec.size() * nbins, 0.0);
    std::vector<char> done(sec.size(), 0);
    for (size_t jsec = 0; jsec < sec.size(); ++jsec) {
        const Section &s = sec[jsec];
        bool merid;
        int npts;
        if (s.imin == s.imax) { npts = s.jmax - s.jmin; merid = true; }
        else if (s.jmin == s.jmax) { npts = s.imax - s.imin; merid = false; }
        else {
            printf(" Section %s is neither zonal nor meridional :(\n We skip this section .\n", s.name.c_str());
            continue;
        }
        if (npts < 1) { printf(" Section %s is empty.\n We skip this section .\n", s.name.c_str()); continue; }
        const size_t n2 = (size_t)npts * npk;
        std::vector<float> eu(npts), de3(n2), e3wa(n2), e3wb(n2), zu(n2), zsa(n2), zsb(n2), zta(n2), ztb(n2);
        std::unique_ptr<Field> e3w;   // (not read with -full)
        if (!lfull) e3w.reset(new Field(zgr.nc, "e3w" + sfx, cn));
        if (merid) {   // :404-460
            Field(hgr, "e2u", cn).yz(s.imin, npts, 1, s.jmin + 1, eu.data());
            if (!lfull) {
                Field(zgr.nc, "e3u" + sfx, cn).yz(s.imin, npts, npk, s.jmin + 1, de3.data());
                e3w->yz(s.imin, npts, npk, s.jmin + 1, e3wa.data());
                e3w->yz(s.imin + 1, npts, npk, s.jmin + 1, e3wb.data());
            }
            U.yz(s.imin, npts, npk, s.jmin + 1, zu.data());
            S.yz(s.imin, npts, npk, s.jmin + 1, zsa.data());
            S.yz(s.imin + 1, npts, npk, s.jmin + 1, zsb.data());
            T.yz(s.imin, npts, npk, s.jmin + 1, zta.data());
            T.yz(s.imin + 1, npts, npk, s.jmin + 1, ztb.data());
        } else {       // :462-555
            Field(hgr, cn.e1v, cn).xz(s.jmin, npts, 1, s.imin, eu.data());   // (the reference starts e1v at imin, the data at imin+1)
            if (!lfull) {
                Field(zgr.nc, "e3v" + sfx, cn).xz(s.jmin, npts, npk, s.imin + 1, de3.data());
                e3w->xz(s.jmin, npts, npk, s.imin + 1, e3wa.data());
                e3w->xz(s.jmin + 1, npts, npk, s.imin + 1, e3wb.data());
            }
            V.xz(s.jmin, npts, npk, s.imin + 1, zu.data());
            S.xz(s.jmin, npts, npk, s.imin + 1, zsa.data());
            S.xz(s.jmin + 1, npts, npk, s.imin + 1, zsb.data());
            T.xz(s.jmin, npts, npk, s.imin + 1, zta.data());
            T.xz(s.jmin + 1, npts, npk, s.imin + 1, ztb.data());
        }
        if (lfull)
            for (int k = 0; k < npk; ++k)
                for (int i = 0; i < npts; ++i) { de3[(size_t)k * npts + i] = e3t1d[k]; e3wa[(size_t)k * npts + i] = e3w1d[k]; e3wb[(size_t)k * npts + i] = e3w1d[k]; }
        // ---- the preparation of :428-460 / :521-555 (REAL(4) expressions, left to right)
        std::vector<double> ddepu(n2 + npts, 0.0);
        std::vector<float> zs(n2), zt(n2), zmask(n2);
        for (int i = 0; i < npts; ++i) ddepu[(size_t)npts + i] = (double)gdept[0];
        for (int k = 1; k < npk; ++k)
            for (int i = 0; i < npts; ++i) {
                const size_t c = (size_t)k * npts + i;
                const float m = e3wa[c] < e3wb[c] ? e3wa[c] : e3wb[c];
                ddepu[c + npts] = ddepu[c] + (double)m;
            }
        const float zsp = merid ? zspu : zspv;
        for (size_t c = 0; c < n2; ++c) {
            if (zu[c] == zsp) zu[c] = 0.f;
            zmask[c] = (zsa[c] == zsps || zsb[c] == zsps) ? 0.f : 1.f;
            volatile float a = zsa[c] + zsb[c];
            a = 0.5f * a;
            zs[c] = a * zmask[c];
            volatile float t = zta[c] + ztb[c];
            t = 0.5f * t;
            zt[c] = merid ? t * zmask[c] : t;
        }
        int nk = npk;   // (the reference leaves nk undefined when no level is all land)
        for (int k = 0; k < npk; ++k) {
            volatile float sum = 0.f;
            for (int i = 0; i < npts; ++i) sum = sum + zs[(size_t)k * npts + i];
            if (sum == 0.f) { nk = k + 1; break; }
        }
        printf(merid ? "  NK =  %d\n" : " JMM nk   %d\n", nk);
        gpu_check(cdfsigtrp_gpu_section(npts, npk, nk, eu.data(), de3.data(), ddepu.data(), gdepw.data(), nullptr, zu.data(), zt.data(),
                                        zs.data(), zmask.data(), mode, refdep, lteos10 ? 1 : 0, dsigma_min, dsigma_max, nbins,
                                        dsigma_lev.data(), nullptr, nullptr, nullptr, nullptr, dtrpbin.data() + jsec * nbins),
                  "cdfsigtrp_gpu_section");
        done[jsec] = 1;
        double tot = 0.0;
        for (int b = 0; b < nbins; ++b) tot += dtrpbin[jsec * nbins + b];
        printf("  Total transport in all bins :%s %.10g\n", s.name.c_str(), tot / 1.e6);
    }
    if (!dsigma_lev.empty() && std::find(done.begin(), done.end(), 1) == done.end()) {   // no section computed: class limits for the files
        const double dlt = (dsigma_max - dsigma_min) / nbins;
        for (int c = 0; c <= nbins; ++c) dsigma_lev[c] = dsigma_min + c * dlt;
    }

    // ---- trpsig.txt (:641-648): formats 9006, 9005, 9004 (:674-676)
    {
        FILE *f = fopen(cf_out.c_str(), "w");
        if (!f) { printf(" ERROR : cannot create %s\n", cf_out.c_str()); stop(99); }
        const size_t ipos = cf_tfil.find("_gridT.nc");
        fprintf(f, "# %s\n", ipos == std::string::npos ? "" : cf_tfil.substr(0, ipos).c_str());
        fprintf(f, "#%9s", " sigma  ");
        for (auto &s : sec) fprintf(f, "  %-12.12s  ", s.name.c_str());
        fprintf(f, "\n");
        for (int b = 0; b < nbins; ++b) {
            fprintf(f, "%9.4f", dsigma_lev[b]);
            for (size_t jsec = 0; jsec < sec.size(); ++jsec) fprintf(f, "%s", e16_7(dtrpbin[jsec * nbins + b]).c_str());
            fprintf(f, "\n");
        }
        fclose(f);
    }
    // ---- one file per section (CreateOutput, :965-1039; the values at :663-668)
    std::vector<double> tim(1, 0.0);
    { const int it = ft.find_var(cn.vtimec); if (it >= 0) ft.read_f64(ft.vars[it], 0, 0, 1, &tim[0]); }
    std::vector<float> lev32(nbins), trp32(nbins), zero(1, 0.f);
    for (int b = 0; b < nbins; ++b) lev32[b] = (float)dsigma_lev[b];
    for (size_t jsec = 0; jsec < sec.size(); ++jsec) {
        const Section &s = sec[jsec];
        const std::string sfxv = s.varname != "none" ? "_" + s.varname : "", pfx = s.longname != "none" ? s.longname + "_" : "";
        std::vector<OutVar> ov(2);
        if (ltemp) {
            ov[0] = {"temp_class", "class of potential temperature", "[]", 0.f, 100.f, 99999.f, "ZT"};
            ov[1] = {"temptrp" + sfxv, pfx + "transport in temperature class", "Sv", -1000.f, 1000.f, 99999.f, "ZT", "temptrp"};
        } else {
            ov[0] = {"sigma_class", "class of potential density", "[]", 0.f, 100.f, 99999.f, "ZT"};
            ov[1] = {"sigtrp" + sfxv, pfx + "transport in sigma class", "Sv", -1000.f, 1000.f, 99999.f, "ZT", "sigtrp"};
        }
        OutFile out;
        out.create(s.name + (ltemp ? "_trptemp.nc" : "_trpsig.nc"), "levels", 1, nbins, ov, cglobal, zero, lev32, tim, ft);
        for (int b = 0; b < nbins; ++b) trp32[b] = (float)(dtrpbin[jsec * nbins + b] / 1.e6);
        out.put(0, 0, lev32.data());
        out.put(1, 0, trp32.data());
        out.w.close();
    }
    gpu_check(cdfgpu_finalize(), "cdfgpu_finalize");
    return 0;
}
