/* A minimal stand-in for the JDK's <jni.h>: only what examples/jni/avocado_b200_jni.c uses, with the JDK's
 * names, types and call shapes (JNI specification, chapter 4), so that the glue can be COMPILED in an image
 * without a JDK (tests/test_packer_cpu.py).  Not a JNI implementation; never linked into anything. */
#ifndef MOCK_JNI_H
#define MOCK_JNI_H
#include <stdint.h>
typedef int32_t jint;
typedef int64_t jlong;
typedef void* jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jarray;
typedef jarray jbyteArray;
typedef int8_t jbyte;
typedef jint jsize;
typedef struct _jfieldID* jfieldID;
struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;
struct JNINativeInterface_ {
  jclass (*GetObjectClass)(JNIEnv*, jobject);
  jfieldID (*GetFieldID)(JNIEnv*, jclass, const char*, const char*);
  jobject (*GetObjectField)(JNIEnv*, jobject, jfieldID);
  jint (*GetIntField)(JNIEnv*, jobject, jfieldID);
  jlong (*GetLongField)(JNIEnv*, jobject, jfieldID);
  void (*SetLongField)(JNIEnv*, jobject, jfieldID, jlong);
  void (*SetIntField)(JNIEnv*, jobject, jfieldID, jint);
  void* (*GetDirectBufferAddress)(JNIEnv*, jobject);
  jlong (*GetDirectBufferCapacity)(JNIEnv*, jobject);
  jstring (*NewStringUTF)(JNIEnv*, const char*);
  void (*GetByteArrayRegion)(JNIEnv*, jbyteArray, jsize, jsize, jbyte*);
  void (*SetByteArrayRegion)(JNIEnv*, jbyteArray, jsize, jsize, const jbyte*);
};
#define JNIEXPORT
#define JNICALL
#endif
