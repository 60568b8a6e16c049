/* include/compat/MultiGrid3D.h -- shim with the public interface of the reference's MultiGrid3D
   (NOCUDA_TESI/POISSON_3D(TESI)/MultiGrid3D.h:6-33) over libmg_b200.so.  Every operator runs on the
   GPU; the host arrays grids3D[l]->h_v / h_f are kept current after each call, which is what the
   reference's main() relies on (N3/Poisson3DSolver.cpp:25-26). */
#ifndef MULTIGRID3D_H
#define MULTIGRID3D_H

#include "Grid3D.h"

class MultiGrid3D
{
	public:
		Grid3D** grids3D;
		int numGrids;
		mg3d_t* engine; // the C-ABI handle behind this object
		mg3b_t* box;    // ... or, for sizeX != sizeY != sizeZ, the non-cubic engine (then engine == 0)

		MultiGrid3D(int finestGridSizeXYZ[], float range[]) { InitGrids(finestGridSizeXYZ, range); }
		~MultiGrid3D()
		{
			for (int i = 0; i < numGrids; i++) delete grids3D[i];
			free(grids3D);
			if (engine) mg3d_destroy(engine);
			if (box) mg3b_destroy(box);
		}
		void InitGrids(int finestGridSizeXYZ[], float range[])
		{
			double r[6];
			for (int i = 0; i < 6; i++) r[i] = range[i];
			engine = 0;
			box = 0;
			if (finestGridSizeXYZ[0] != finestGridSizeXYZ[1] || finestGridSizeXYZ[0] != finestGridSizeXYZ[2]) {
				/* the hierarchy of N3/MultiGrid3D.cpp:19-47 as written: numGrids from the smallest size, every size halved per
				   level -- only Grid3D's asserts (N3/Grid3D.cpp:10-11) keep the reference itself from building it */
#ifdef MG_COMPAT_CUDA_TESI
				fprintf(stderr, "MultiGrid3D: non-cubic grids are served by the NOCUDA_TESI face of the shim (and by mg3b_* directly)\n");
				abort();
#else
				MG_CHECK(mg3b_create(&box, finestGridSizeXYZ, r, MG_F32, MG_REF_COMPAT));
				numGrids = mg3b_num_levels(box);
				grids3D = (Grid3D**)malloc(numGrids * sizeof(Grid3D*));
				for (int l = 0; l < numGrids; l++) {
					int s[3];
					MG_CHECK(mg3b_level_size(box, l, s));
					grids3D[l] = new Grid3D(s, range, box, l);
				}
				return;
#endif
			}
			MG_CHECK(mg3d_create(&engine, finestGridSizeXYZ, r, MG_F32, MG_REF_COMPAT));
			numGrids = mg3d_num_levels(engine);
			grids3D = (Grid3D**)malloc(numGrids * sizeof(Grid3D*));
			for (int l = 0; l < numGrids; l++) {
				int n = mg3d_level_size(engine, l);
				int s[3] = {n, n, n};
				grids3D[l] = new Grid3D(s, range, engine, l);
			}
		}

		void Relax(Grid3D* curGrid, int ncycles)
		{
			int l = level_of(curGrid);
#ifndef MG_COMPAT_CUDA_TESI
			if (box) {
				curGrid->push(box, l);
				MG_CHECK(mg3b_relax(box, l, ncycles));
				curGrid->pull(box, l);
				return;
			}
#endif
			curGrid->push(engine, l);
			MG_CHECK(mg3d_relax(engine, l, ncycles));
			curGrid->pull(engine, l);
		}
#ifdef MG_COMPAT_CUDA_TESI
		/* CUDA_TESI face (C3/MultiGrid3D.h:16-24): the arrays AND the size triplets are device pointers; like the reference's own
		   wrappers (C3/MultiGrid3D.cu:68-72) the triplets are copied back before the launch */
		void Restrict(float* d_fine, int d_fsizeXYZ[], float* d_coarse, int d_csizeXYZ[])
		{
			int fs[3], cs[3];
			dsize(d_fsizeXYZ, fs); dsize(d_csizeXYZ, cs);
			MG_CHECK(mg3d_restrict_device(engine, d_fine, fs, d_coarse, cs));
		}
		void Interpolate(float* d_fine, int d_fsizeXYZ[], float* d_coarse, int d_csizeXYZ[])
		{
			int fs[3], cs[3];
			dsize(d_fsizeXYZ, fs); dsize(d_csizeXYZ, cs);
			MG_CHECK(mg3d_interpolate_device(engine, d_fine, fs, d_coarse, cs));
		}
		void ApplyCorrection(float* d_fine, int d_fsizeXYZ[], float* d_error, int d_esizeXYZ[])
		{
			int fs[3], es[3];
			dsize(d_fsizeXYZ, fs); dsize(d_esizeXYZ, es);
			MG_CHECK(mg3d_apply_correction_device(engine, d_fine, fs, d_error, es));
		}
		void Set(float* d_v, int d_sizeXYZ[], float value, bool modifyBorder)
		{
			int s[3];
			dsize(d_sizeXYZ, s);
			MG_CHECK(mg3d_set_device(engine, d_v, s, value, modifyBorder));
		}
		/* the index-map probe of the twin (C3/MultiGrid3D.cu:101-118,702-720): v(x,y,z) = x + y + z */
		void SetTESTTEST(float* d_v, int d_sizeXYZ[], float, bool)
		{
			int s[3];
			dsize(d_sizeXYZ, s);
			size_t tot = (size_t)s[0] * s[1] * s[2];
			float* h = (float*)malloc(tot * sizeof(float));
			for (int z = 0; z < s[2]; z++)
				for (int y = 0; y < s[1]; y++)
					for (int x = 0; x < s[0]; x++) h[x + (size_t)y * s[0] + (size_t)z * s[0] * s[1]] = (float)(x + y + z);
			MG_CUDA_CHECK(cudaMemcpy(d_v, h, tot * sizeof(float), cudaMemcpyHostToDevice));
			free(h);
		}
		float* CalculateResidual(Grid3D* curGrid) // caller owns the returned DEVICE buffer, as in the twin (C3/MultiGrid3D.cu:220-222)
		{
			int l = level_of(curGrid);
			curGrid->push(engine, l);
			float* d_r = 0;
			MG_CUDA_CHECK(cudaMalloc((void**)&d_r, (size_t)curGrid->sizeX * curGrid->sizeY * curGrid->sizeZ * sizeof(float)));
			MG_CHECK(mg3d_residual_device(engine, l, d_r));
			return d_r;
		}
#else
		void Restrict(float* fine, int fsizeXYZ[], float* coarse, int csizeXYZ[])
		{
			if (box) MG_CHECK(mg3b_restrict_host(box, fine, fsizeXYZ, coarse, csizeXYZ));
			else MG_CHECK(mg3d_restrict_host(engine, fine, fsizeXYZ, coarse, csizeXYZ));
		}
		void Interpolate(float* fine, int fsizeXYZ[], float* coarse, int csizeXYZ[])
		{
			if (box) MG_CHECK(mg3b_interpolate_host(box, fine, fsizeXYZ, coarse, csizeXYZ));
			else MG_CHECK(mg3d_interpolate_host(engine, fine, fsizeXYZ, coarse, csizeXYZ));
		}
		void setToValue(float* grid, int sizeXYZ[], float value, bool modifyBoundaries)
		{
			if (box) MG_CHECK(mg3b_set_to_value_host(box, grid, sizeXYZ, value, modifyBoundaries));
			else MG_CHECK(mg3d_set_to_value_host(engine, grid, sizeXYZ, value, modifyBoundaries));
		}
		float* CalculateResidual(Grid3D* fine) // caller owns the returned buffer, as in the reference
		{
			int l = level_of(fine);
			float* r = (float*)malloc((size_t)fine->sizeX * fine->sizeY * fine->sizeZ * sizeof(float));
			if (box) {
				fine->push(box, l);
				MG_CHECK(mg3b_residual(box, l, r));
				return r;
			}
			fine->push(engine, l);
			MG_CHECK(mg3d_residual(engine, l, r));
			return r;
		}
		void ApplyCorrection(float* fine, int fsizeXYZ[], float* error, int esizeXYZ[])
		{
			if (box) MG_CHECK(mg3b_apply_correction_host(box, fine, fsizeXYZ, error, esizeXYZ));
			else MG_CHECK(mg3d_apply_correction_host(engine, fine, fsizeXYZ, error, esizeXYZ));
		}
#endif

		void VCycle(int gridID, int v1, int v2)
		{
			push_all();
#ifndef MG_COMPAT_CUDA_TESI
			if (box) MG_CHECK(mg3b_vcycle(box, gridID, v1, v2));
			else
#endif
			MG_CHECK(mg3d_vcycle(engine, gridID, v1, v2));
			pull_all();
		}
		void FullMultiGridVCycle(int gridID, int v0, int v1, int v2)
		{
			push_all();
#ifndef MG_COMPAT_CUDA_TESI
			if (box) MG_CHECK(mg3b_fmg(box, gridID, v0, v1, v2));
			else
#endif
			MG_CHECK(mg3d_fmg(engine, gridID, v0, v1, v2));
			pull_all();
		}

		void PrintGrid(int gridID) { grids3D[gridID]->PrintGrid_v(mg_compat_open_log("log/log_v.txt")); }
		void PrintAllGrids_v() { int fd = mg_compat_open_log("log/log_v.txt"); for (int i = 0; i < numGrids; i++) grids3D[i]->PrintGrid_v(fd); }
		void PrintAllGrids_f() { int fd = mg_compat_open_log("log/log_f.txt"); for (int i = 0; i < numGrids; i++) grids3D[i]->PrintGrid_f(fd); }
		void PrintDiff() { grids3D[0]->PrintDiff(mg_compat_open_log("log/diff.txt")); }

	private:
#ifdef MG_COMPAT_CUDA_TESI
		static void dsize(const int* d_size, int out[3]) { MG_CUDA_CHECK(cudaMemcpy(out, d_size, 3 * sizeof(int), cudaMemcpyDeviceToHost)); }
#endif
		int level_of(Grid3D* g)
		{
			for (int l = 0; l < numGrids; l++) if (grids3D[l] == g) return l;
			fprintf(stderr, "MultiGrid3D: grid does not belong to this hierarchy\n");
			abort();
		}
		void push_all()
		{
			for (int l = 0; l < numGrids; l++) {
#ifndef MG_COMPAT_CUDA_TESI
				if (box) { grids3D[l]->push(box, l); continue; }
#endif
				grids3D[l]->push(engine, l);
			}
		}
		void pull_all()
		{
			for (int l = 0; l < numGrids; l++) {
#ifndef MG_COMPAT_CUDA_TESI
				if (box) { grids3D[l]->pull(box, l); continue; }
#endif
				grids3D[l]->pull(engine, l);
			}
		}
};
#endif
