// shud_io.cu - host-side ingest and checkpoint around the device path.
//
// (1) Binary mesh container.  The reference builds Model_Data from ~10 text files with strtold per field into
//     double** rows (src/classes/TabularData.cpp:27-55, src/ModelData/MD_readin.cpp:192-236): minutes for an
//     8M-cell mesh.  shud_b200_mesh_save writes the shud_mesh SoA a host has exported once (INTEGRATION.md
//     section 1) as one file: header, then every array 64-byte aligned; shud_b200_mesh_load reads it back with a
//     single read() into one block and points a shud_mesh into it - ingest at file-system speed.
// (2) Checkpoint in the reference's own initial-condition format (Model_Data::PrintInit,
//     src/ModelData/MD_update.cpp:268-299: "<prj>.cfg.ic.update", %lf columns): shud_b200_write_ic takes the
//     state straight from the device (device order -> reference order, plus the land-surface buckets when the
//     land step runs on the device), shud_b200_format_ic is the host-only formatter (byte-identical to the
//     reference's writer, tests/test_io_cpu.py).
// No CUDA kernels here; compiled into the same library.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "shud_b200.h"

namespace {

constexpr char kMagic[8] = {'S', 'H', 'U', 'D', 'B', '2', '0', '0'};
constexpr uint32_t kVersion = 1;

struct Field {
    const char *name;
    int is_int;     // 0 double, 1 int32
    size_t offset;  // of the pointer inside shud_mesh
    int dim;        // 0 Ne, 1 3*Ne, 2 Nr, 3 Ns, 4 Nl, 5 Nl+1, 6 nbathy (lake_bathy_ptr[Nl])
    int optional;
};
#define FD(n, d) {#n, 0, offsetof(shud_mesh, n), d, 0}
#define FI(n, d) {#n, 1, offsetof(shud_mesh, n), d, 0}
const Field kFields[] = {
    FD(area, 0), FD(z_surf, 0), FD(z_bottom, 0), FD(depression, 0), FD(AquiferDepth, 0), FD(Sy, 0), FD(infD, 0),
    FD(infKsatV, 0), FD(macKsatV, 0), FD(hAreaF, 0), FD(ThetaS, 0), FD(ThetaR, 0), FD(ThetaFC, 0), FD(Beta, 0),
    FD(KsatH, 0), FD(KsatV, 0), FD(macKsatH, 0), FD(macD, 0), FD(geo_vAreaF, 0), FD(VegFrac, 0), FD(ImpAF, 0),
    FD(WetlandLevel, 0), FD(RootReachLevel, 0), FD(Rough, 0), FD(QSS, 0), FD(edge, 1), FD(Dist2Nabor, 1),
    FD(Dist2Edge, 1), FD(avgRough, 1), FI(nabr, 1), FI(lakenabr, 1), FI(iLake, 0), FI(iBC, 0), FI(iSS, 0),
    {"x", 0, offsetof(shud_mesh, x), 0, 1}, {"y", 0, offsetof(shud_mesh, y), 0, 1},
    FD(riv_Length, 2), FD(riv_BedSlope, 2), FD(riv_depth, 2), FD(riv_BottomWidth, 2), FD(riv_bankslope, 2),
    FD(riv_avgRough, 2), FD(riv_Dist2DownStream, 2), FD(riv_KsatH, 2), FD(riv_BedThick, 2), FD(riv_zbank, 2),
    FI(riv_down, 2), FI(riv_BC, 2), FI(riv_toLake, 2), FI(seg_iEle, 3), FI(seg_iRiv, 3), FD(seg_length, 3),
    FD(seg_Cwr, 3), FD(lake_zmin, 4), FI(lake_NumEleLake, 4), FI(lake_bathy_ptr, 5), FD(lake_bathy_yi, 6),
    FD(lake_bathy_ai, 6)};
constexpr int kNF = sizeof(kFields) / sizeof(kFields[0]);

size_t dim_len(const shud_mesh *m, int dim, size_t nbathy) {
    switch (dim) {
        case 0: return (size_t)m->Ne;
        case 1: return 3 * (size_t)m->Ne;
        case 2: return (size_t)m->Nr;
        case 3: return (size_t)m->Ns;
        case 4: return (size_t)m->Nl;
        case 5: return m->Nl > 0 ? (size_t)m->Nl + 1 : 0;
        default: return nbathy;
    }
}
const void *get_ptr(const shud_mesh *m, const Field &f) {
    const void *p;
    memcpy(&p, (const char *)m + f.offset, sizeof(p));
    return p;
}
struct Header {
    char magic[8];
    uint32_t version, nfields;
    int32_t Ne, Nr, Ns, Nl, close_boundary, lakeon;
    uint64_t nbathy, payload_bytes;
};
struct Entry {
    char name[24];
    uint32_t is_int, present;
    uint64_t count, offset;  // offset of the array from the start of the payload (64-byte aligned)
};
size_t align64(size_t x) { return (x + 63) & ~(size_t)63; }

}  // namespace

extern "C" {

int shud_b200_mesh_save(const char *path, const shud_mesh *m) {
    if (!path || !m || m->Ne <= 0) return SHUD_ERR_ARG;
    const size_t nbathy = (m->Nl > 0 && m->lake_bathy_ptr) ? (size_t)m->lake_bathy_ptr[m->Nl] : 0;
    Header h = {};
    memcpy(h.magic, kMagic, 8);
    h.version = kVersion; h.nfields = kNF;
    h.Ne = m->Ne; h.Nr = m->Nr; h.Ns = m->Ns; h.Nl = m->Nl; h.close_boundary = m->close_boundary; h.lakeon = m->lakeon;
    h.nbathy = nbathy;
    std::vector<Entry> ent(kNF);
    size_t off = 0;
    for (int k = 0; k < kNF; k++) {
        const Field &f = kFields[k];
        Entry &e = ent[k];
        memset(&e, 0, sizeof(e));
        strncpy(e.name, f.name, sizeof(e.name) - 1);
        e.is_int = f.is_int;
        const void *p = get_ptr(m, f);
        e.count = dim_len(m, f.dim, nbathy);
        e.present = (p && e.count) ? 1 : 0;
        if (!p && e.count && !f.optional) return SHUD_ERR_ARG;
        e.offset = off;
        if (e.present) off = align64(off + e.count * (f.is_int ? 4 : 8));
    }
    h.payload_bytes = off;
    FILE *fp = fopen(path, "wb");
    if (!fp) return SHUD_ERR_ARG;
    bool ok = fwrite(&h, sizeof(h), 1, fp) == 1 && fwrite(ent.data(), sizeof(Entry), kNF, fp) == (size_t)kNF;
    const size_t head = sizeof(h) + sizeof(Entry) * kNF, pad0 = align64(head) - head;
    static const char zeros[64] = {0};
    ok = ok && (pad0 == 0 || fwrite(zeros, 1, pad0, fp) == pad0);
    for (int k = 0; k < kNF && ok; k++) {
        if (!ent[k].present) continue;
        const size_t nb = ent[k].count * (ent[k].is_int ? 4 : 8), pad = align64(nb) - nb;
        ok = fwrite(get_ptr(m, kFields[k]), 1, nb, fp) == nb && (pad == 0 || fwrite(zeros, 1, pad, fp) == pad);
    }
    ok = (fclose(fp) == 0) && ok;
    return ok ? SHUD_OK : SHUD_ERR_CUDA;
}

int shud_b200_mesh_load(const char *path, shud_mesh *out, void **block) {
    if (!path || !out || !block) return SHUD_ERR_ARG;
    *block = nullptr;
    FILE *fp = fopen(path, "rb");
    if (!fp) return SHUD_ERR_ARG;
    Header h;
    std::vector<Entry> ent;
    bool ok = fread(&h, sizeof(h), 1, fp) == 1 && memcmp(h.magic, kMagic, 8) == 0 && h.version == kVersion &&
              h.nfields == (uint32_t)kNF;
    if (ok) {
        ent.resize(kNF);
        ok = fread(ent.data(), sizeof(Entry), kNF, fp) == (size_t)kNF;
    }
    char *buf = nullptr;
    if (ok) {
        // the payload must fit in what the file actually holds: a crafted header must not size the allocation
        const size_t head = sizeof(h) + sizeof(Entry) * kNF;
        ok = fseek(fp, 0, SEEK_END) == 0;
        const long fsize = ok ? ftell(fp) : -1;
        ok = ok && fsize >= 0 && (uint64_t)fsize >= align64(head) && h.payload_bytes <= (uint64_t)fsize - align64(head);
    }
    if (ok) {
        const size_t head = sizeof(h) + sizeof(Entry) * kNF;
        ok = fseek(fp, (long)align64(head), SEEK_SET) == 0;
        if (ok && posix_memalign((void **)&buf, 64, h.payload_bytes ? h.payload_bytes : 64) != 0) { buf = nullptr; ok = false; }
        if (ok && h.payload_bytes) ok = fread(buf, 1, h.payload_bytes, fp) == h.payload_bytes;  // one read
    }
    fclose(fp);
    if (!ok) { free(buf); return SHUD_ERR_ARG; }
    memset(out, 0, sizeof(*out));
    out->Ne = h.Ne; out->Nr = h.Nr; out->Ns = h.Ns; out->Nl = h.Nl; out->close_boundary = h.close_boundary; out->lakeon = h.lakeon;
    for (int k = 0; k < kNF; k++) {
        if (strncmp(ent[k].name, kFields[k].name, sizeof(ent[k].name)) != 0 || ent[k].is_int != (uint32_t)kFields[k].is_int ||
            ent[k].count != dim_len(out, kFields[k].dim, h.nbathy) ||
            // offsets come from the file: no wrap-around, inside the payload, aligned for the element type
            (ent[k].present && (ent[k].offset > h.payload_bytes || (ent[k].offset & 63) != 0 ||
                                ent[k].count > (h.payload_bytes - ent[k].offset) / (ent[k].is_int ? 4 : 8)))) {
            free(buf);
            return SHUD_ERR_ARG;
        }
        const void *p = ent[k].present ? buf + ent[k].offset : nullptr;
        memcpy((char *)out + kFields[k].offset, &p, sizeof(p));
    }
    *block = buf;
    return SHUD_OK;
}

void shud_b200_mesh_free(void *block) { free(block); }

// Model_Data::PrintInit, src/ModelData/MD_update.cpp:268-299 (the UpdateICStep gate stays with the caller)
int shud_b200_format_ic(const char *path, double t, int32_t Ne, int32_t Nr, int32_t Nl, const double *yEleIS,
                        const double *yEleSnow, const double *y) {
    if (!path || !y || Ne <= 0 || Nr < 0 || Nl < 0) return SHUD_ERR_ARG;
    FILE *fp = fopen(path, "w");
    if (!fp) return SHUD_ERR_ARG;
    const double *sf = y, *us = y + Ne, *gw = y + 2 * (size_t)Ne, *riv = y + 3 * (size_t)Ne, *lake = riv + Nr;
    fprintf(fp, "%d\t %d \t%lf\n", Ne, 6, t);
    fprintf(fp, "%s\t%s\t%s\t%s\t%s\t%s\n", "Index", "Canopy", "Snow", "Surface", "Unsat", "GW");
    for (int i = 0; i < Ne; i++)
        fprintf(fp, "%d\t%lf\t%lf\t%lf\t%lf\t%lf\n", i + 1, yEleIS ? yEleIS[i] : 0., yEleSnow ? yEleSnow[i] : 0., sf[i], us[i], gw[i]);
    fprintf(fp, "%d\t%d\n", Nr, 2);
    fprintf(fp, "%s\t%s\n", "Index", "Stage");
    for (int i = 0; i < Nr; i++) fprintf(fp, "%d\t%lf\n", i + 1, riv[i]);
    if (Nl > 0) {
        fprintf(fp, "%d\t%d\n", Nl, 2);
        fprintf(fp, "%s\t%s\n", "Index", "LakeStage");
        for (int i = 0; i < Nl; i++) fprintf(fp, "%d\t%lf\n", i + 1, lake[i]);
    }
    return fclose(fp) == 0 ? SHUD_OK : SHUD_ERR_CUDA;
}

// The reverse of shud_b200_format_ic: read "<prj>.cfg.ic" / ".cfg.ic.update" (the %lf text of PrintInit) back into the
// blocked state vector and the two land-surface buckets.  Sizes must match the mesh the file was written for.
int shud_b200_read_ic(const char *path, int32_t Ne, int32_t Nr, int32_t Nl, double *t, double *yEleIS, double *yEleSnow,
                      double *y) {
    if (!path || !y || Ne <= 0 || Nr < 0 || Nl < 0) return SHUD_ERR_ARG;
    FILE *fp = fopen(path, "r");
    if (!fp) return SHUD_ERR_ARG;
    char line[512];
    int n = 0, ncol = 0, idx = 0;
    double tt = 0.;
    bool ok = fgets(line, sizeof line, fp) && sscanf(line, "%d %d %lf", &n, &ncol, &tt) == 3 && n == Ne && ncol == 6 &&
              fgets(line, sizeof line, fp);  // column names
    for (int i = 0; ok && i < Ne; i++) {
        double is, sn, sf, us, gw;
        ok = fgets(line, sizeof line, fp) && sscanf(line, "%d %lf %lf %lf %lf %lf", &idx, &is, &sn, &sf, &us, &gw) == 6 &&
             idx == i + 1;
        if (ok) {
            if (yEleIS) yEleIS[i] = is;
            if (yEleSnow) yEleSnow[i] = sn;
            y[i] = sf; y[(size_t)Ne + i] = us; y[2 * (size_t)Ne + i] = gw;
        }
    }
    ok = ok && fgets(line, sizeof line, fp) && sscanf(line, "%d %d", &n, &ncol) == 2 && n == Nr && ncol == 2 &&
         fgets(line, sizeof line, fp);
    for (int i = 0; ok && i < Nr; i++) {
        double v;
        ok = fgets(line, sizeof line, fp) && sscanf(line, "%d %lf", &idx, &v) == 2 && idx == i + 1;
        if (ok) y[3 * (size_t)Ne + i] = v;
    }
    if (ok && Nl > 0) {
        ok = fgets(line, sizeof line, fp) && sscanf(line, "%d %d", &n, &ncol) == 2 && n == Nl && ncol == 2 &&
             fgets(line, sizeof line, fp);
        for (int i = 0; ok && i < Nl; i++) {
            double v;
            ok = fgets(line, sizeof line, fp) && sscanf(line, "%d %lf", &idx, &v) == 2 && idx == i + 1;
            if (ok) y[3 * (size_t)Ne + Nr + i] = v;
        }
    }
    fclose(fp);
    if (ok && t) *t = tt;
    return ok ? SHUD_OK : SHUD_ERR_ARG;
}

}  // extern "C"
