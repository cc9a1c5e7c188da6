// nc3dump -- prints the header of a NetCDF classic file and a checksum of every numeric variable as JSON lines.
// Test utility for the self-contained reader/writer (tests/test_host_nc3.py); no GPU involved.
#include "nc3.hpp"

int main(int argc, char **argv)
{
    if (argc < 2) { printf("usage: nc3dump file.nc\n"); return 99; }
    nc3::Reader r;
    if (!r.open(argv[1])) { printf("{\"error\": \"%s\"}\n", r.err.c_str()); return 98; }
    printf("{\"version\": %d, \"numrecs\": %llu, \"dims\": {", r.version, (unsigned long long)r.numrecs);
    for (size_t i = 0; i < r.dims.size(); ++i) printf("%s\"%s\": %llu", i ? ", " : "", r.dims[i].name.c_str(), (unsigned long long)r.dims[i].len);
    printf("}, \"vars\": {");
    bool first = true;
    for (auto &v : r.vars) {
        if (v.type == nc3::NC_CHAR) continue;
        const uint64_t nrec = v.isrec ? r.numrecs : 1;
        double sum = 0.0, asum = 0.0;
        std::vector<double> buf(v.nelem_per_rec);
        for (uint64_t rec = 0; rec < nrec; ++rec) {
            if (!r.read_f64(v, (long)rec, 0, v.nelem_per_rec, buf.data())) { printf("{\"error\": \"%s\"}\n", r.err.c_str()); return 98; }
            for (double x : buf) { sum += x; asum += x < 0 ? -x : x; }
        }
        printf("%s\"%s\": {\"type\": %d, \"nelem\": %llu, \"sum\": %.17g, \"abssum\": %.17g, \"spval\": %.9g}", first ? "" : ", ",
               v.name.c_str(), v.type, (unsigned long long)(v.nelem_per_rec * nrec), sum, asum, (double)r.spval(v));
        first = false;
    }
    printf("}}\n");
    return 0;
}
