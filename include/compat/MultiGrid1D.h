/* include/compat/MultiGrid1D.h -- shim with the public interface of the reference's MultiGrid1D
   (NOCUDA_TESI/EQUAZIONE 1D/MultiGrid1D.h:6-31) over libmg_b200.so. */
#ifndef MULTIGRID1D_H
#define MULTIGRID1D_H

#include "Grid1D.h"

class MultiGrid1D
{
	public:
		Grid1D** grids1D;
		int numGrids;
		mg1d_t* engine;

		MultiGrid1D(int finestGridSize, float range[]) { InitGrids(finestGridSize, range); }
		~MultiGrid1D()
		{
			for (int i = 0; i < numGrids; i++) delete grids1D[i];
			free(grids1D);
			mg1d_destroy(engine);
		}
		void InitGrids(int finestGridSize, float range[])
		{
			double r[2] = {range[0], range[1]};
			MG_CHECK(mg1d_create(&engine, finestGridSize, r, MG_F32, MG_REF_COMPAT));
			numGrids = mg1d_num_levels(engine);
			grids1D = (Grid1D**)malloc(numGrids * sizeof(Grid1D*));
			for (int l = 0; l < numGrids; l++) grids1D[l] = new Grid1D(mg1d_level_size(engine, l), range, engine, l);
		}

		void Relax(Grid1D* curGrid, int ncycles)
		{
			int l = level_of(curGrid);
			curGrid->push(engine, l);
			MG_CHECK(mg1d_relax(engine, l, ncycles));
			curGrid->pull(engine, l);
		}
#ifdef MG_COMPAT_CUDA_TESI
		/* CUDA_TESI face (C1/MultiGrid1D.h:16-23): the operands are DEVICE arrays */
		void Restrict(float* fine, int fsize, float* coarse, int csize) { MG_CHECK(mg1d_restrict_device(engine, fine, fsize, coarse, csize)); }
		void Interpolate(float* fine, int fsize, float* coarse, int csize) { MG_CHECK(mg1d_interpolate_device(engine, fine, fsize, coarse, csize)); }
		void ApplyCorrection(float* fine, int fineSize, float* error, int errorSize) { MG_CHECK(mg1d_apply_correction_device(engine, fine, fineSize, error, errorSize)); }
		void Set(float* d_v, int sizeX, float value, bool modifyBoundaries) { MG_CHECK(mg1d_set_device(engine, d_v, sizeX, value, modifyBoundaries)); }
		float* CalculateResidual(Grid1D* fine) // caller-owned DEVICE array (C1/MultiGrid1D.cu:86-104)
		{
			int l = level_of(fine);
			float* d_r = 0;
			MG_CUDA_CHECK(cudaMalloc((void**)&d_r, (size_t)fine->sizeX * sizeof(float)));
			MG_CHECK(mg1d_residual_device(engine, l, d_r));
			return d_r;
		}
#else
		void Restrict(float* fine, int fsize, float* coarse, int csize) { MG_CHECK(mg1d_restrict_host(engine, fine, fsize, coarse, csize)); }
		void Interpolate(float* fine, int fsize, float* coarse, int csize) { MG_CHECK(mg1d_interpolate_host(engine, fine, fsize, coarse, csize)); }
		float* CalculateResidual(Grid1D* fine)
		{
			int l = level_of(fine);
			fine->push(engine, l);
			float* r = (float*)malloc((size_t)fine->sizeX * sizeof(float));
			MG_CHECK(mg1d_residual(engine, l, r));
			return r;
		}
		void ApplyCorrection(float* fine, int fineSize, float* error, int errorSize) { MG_CHECK(mg1d_apply_correction_host(engine, fine, fineSize, error, errorSize)); }
#endif
		void setToValue(float* grid, int sizeX, float value, bool modifyBoundaries) { MG_CHECK(mg1d_set_to_value_host(engine, grid, sizeX, value, modifyBoundaries)); }

		void VCycle(int gridID, int v1, int v2)
		{
			push_all();
			MG_CHECK(mg1d_vcycle(engine, gridID, v1, v2));
			pull_all();
		}
		void FullMultiGridVCycle(int gridID, int v0, int v1, int v2)
		{
			push_all();
			MG_CHECK(mg1d_fmg(engine, gridID, v0, v1, v2));
			pull_all();
		}

		void PrintDiff() { grids1D[0]->PrintDiffApproxReal(mg_compat_open_log("log/diff.txt")); }
		void PrintGrid(int gridID) { grids1D[gridID]->PrintGrid_v(mg_compat_open_log("log/log_v.txt")); }
		void PrintAllGrids_v() { int fd = mg_compat_open_log("log/log_v.txt"); for (int i = 0; i < numGrids; i++) grids1D[i]->PrintGrid_v(fd); }
		void PrintAllGrids_f() { int fd = mg_compat_open_log("log/log_f.txt"); for (int i = 0; i < numGrids; i++) grids1D[i]->PrintGrid_f(fd); }

	private:
		int level_of(Grid1D* g)
		{
			for (int l = 0; l < numGrids; l++) if (grids1D[l] == g) return l;
			fprintf(stderr, "MultiGrid1D: grid does not belong to this hierarchy\n");
			abort();
		}
		void push_all() { for (int l = 0; l < numGrids; l++) grids1D[l]->push(engine, l); }
		void pull_all() { for (int l = 0; l < numGrids; l++) grids1D[l]->pull(engine, l); }
};
#endif
