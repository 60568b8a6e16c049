/* include/compat/Grid2D.h -- shim with the public interface of the reference's Grid2D
   (NOCUDA_TESI/PDE Lyapunov 2D/Grid2D.h:4-33).  Same include guard as the reference header. */
#ifndef GRID2D_H
#define GRID2D_H

#include "mg_compat_common.h"

class Grid2D
{
	public:
		float* h_v;
		float* h_f;

		int sizeX;
		int sizeY;
		int* sizeXY;

		float h_x;
		float h_y;

		float x_a;
		float x_b;
		float y_a;
		float y_b;

#ifdef MG_COMPAT_CUDA_TESI
		/* CUDA_TESI face (CUDA_TESI/CUDA Lyapunov 2D/Grid2D.h:7,19-22): scalar size and the two fields as pitched DEVICE arrays.
		   They ARE the engine's storage of the level (its 2D layout is pitched too): nothing is mirrored, a caller may hand
		   them to the operators or write into them.  NULL for a Grid2D constructed on its own. */
		int size;
		float* d_v;
		float* d_f;
		size_t d_pitchByte;
		int d_pitch;
#endif

		Grid2D(int sizeXY_[], float range[]) { setup(sizeXY_, range); InitV(); InitF(); }
		Grid2D(int sizeXY_[], float range[], mg2d_t* mg, int level) { setup(sizeXY_, range); attach(mg, level); pull(mg, level); }
		~Grid2D() { free(h_v); free(h_f); free(sizeXY); }

		void InitV() { fetch(MG_FIELD_V); }
		void InitF() { fetch(MG_FIELD_F); }

		void pull(mg2d_t* mg, int level)
		{
			MG_CHECK(mg2d_get_field(mg, level, MG_FIELD_V, h_v));
			MG_CHECK(mg2d_get_field(mg, level, MG_FIELD_F, h_f));
		}
		void push(mg2d_t* mg, int level) const
		{
#ifndef MG_COMPAT_CUDA_TESI /* (CUDA_TESI face: the device arrays are the engine's own, the host arrays only feed the dumps) */
			MG_CHECK(mg2d_set_field(mg, level, MG_FIELD_V, h_v));
			MG_CHECK(mg2d_set_field(mg, level, MG_FIELD_F, h_f));
#else
			(void)mg; (void)level;
#endif
		}
		void attach(mg2d_t* mg, int level)
		{
#ifdef MG_COMPAT_CUDA_TESI
			void *pv = 0, *pf = 0;
			MG_CHECK(mg2d_level_device_ptr(mg, level, MG_FIELD_V, &pv, &d_pitch));
			MG_CHECK(mg2d_level_device_ptr(mg, level, MG_FIELD_F, &pf, &d_pitch));
			d_v = (float*)pv; d_f = (float*)pf;
			d_pitchByte = (size_t)d_pitch * sizeof(float);
#else
			(void)mg; (void)level;
#endif
		}

		void PrintDiffApproxReal(int diff_fd)
		{
			char line[200];
			for (int y = 0; y < sizeY; y++)
				for (int x = 0; x < sizeX; x++) {
					float xj = x_a + x * h_x, yi = y_a + y * h_y;
					float exact = 2 * xj * xj - 4 * xj * yi + 2 * yi * yi;
					snprintf(line, sizeof line, "yi: %f xj: %f diff: %f\n", yi, xj, h_v[x + y * sizeX] - exact);
					mg_compat_write(diff_fd, line);
				}
		}
		void PrintGrid_v(int logfd) { dump(logfd, h_v); }
		void PrintGrid_f(int logfd) { dump(logfd, h_f); }
		void PrintResidual(int, float*, int) {}
		void PrintMeanAbsoluteError() { printf("MeanAbsoluteError: %f\n", MeanAbsError()); } // CUDA twin, C2/Grid2D.cu:123-154
		double MeanAbsError() const
		{
			double tot = 0;
			long cnt = 0;
			for (int y = 1; y < sizeY - 1; y++)
				for (int x = 1; x < sizeX - 1; x++) {
					float xj = x_a + x * h_x, yi = y_a + y * h_y;
					float exact = 2 * xj * xj - 4 * xj * yi + 2 * yi * yi;
					tot += fabs(h_v[x + y * sizeX] - exact);
					cnt++;
				}
			return cnt ? tot / cnt : 0.0;
		}

	private:
		float range_[4];
		void setup(int s[], float range[])
		{
			sizeX = s[0]; sizeY = s[1];
			sizeXY = (int*)malloc(2 * sizeof(int));
			sizeXY[0] = sizeX; sizeXY[1] = sizeY;
			for (int i = 0; i < 4; i++) range_[i] = range[i];
			x_a = range[0]; x_b = range[1]; y_a = range[2]; y_b = range[3];
			h_x = (x_b - x_a) / (float)(sizeX - 1);
			h_y = (y_b - y_a) / (float)(sizeY - 1);
			h_v = (float*)malloc((size_t)sizeX * sizeY * sizeof(float));
			h_f = (float*)malloc((size_t)sizeX * sizeY * sizeof(float));
#ifdef MG_COMPAT_CUDA_TESI
			size = sizeX; d_v = d_f = 0; d_pitchByte = 0; d_pitch = 0;
#endif
		}
		void fetch(int field)
		{
			double r[4], A[4] = {0, 0, 0, 0};
			for (int i = 0; i < 4; i++) r[i] = range_[i];
			mg2d_t* mg = 0;
			MG_CHECK(mg2d_create(&mg, sizeXY, r, A, 0, MG_F32)); // InitV/InitF do not depend on A, alfa
			MG_CHECK(mg2d_get_field(mg, 0, field, field == MG_FIELD_V ? h_v : h_f));
			mg2d_destroy(mg);
		}
		void dump(int logfd, const float* a)
		{
			char line[200];
			for (int y = 0; y < sizeY; y++)
				for (int x = 0; x < sizeX; x++) {
					snprintf(line, sizeof line, "posY: %d posX: %d value: %f\n", y, x, a[x + y * sizeX]);
					mg_compat_write(logfd, line);
				}
		}
};
#endif
